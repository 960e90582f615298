"""CPU, dev container only (skipped where /root/reference is absent, e.g. on the GPU box): the reference's OWN, unmodified
Python package (shredword/cbase.py + trainer.py) and its OWN pytest file (test/test_bpe.py) against our libraries
(SURVEY.md section 7 step 2's gate).  Nothing is copied: a temporary directory holds symlinks to the reference's .py files and,
where the reference's loader looks for it (cbase.py:4-19: <pkg>/lib/libtrainer*.so), a symlink to our library.
  * the CUDA library: `import shredword` must resolve all 21 symbols the reference binds eagerly (cbase.py:50-71); creating a
    trainer is not possible without a GPU;
  * the hostsim library (the product's ABI + host control code linked to a CPU stand-in of the device engine): the reference's
    test_bpe.py runs unmodified, end to end, through the reference's own ctypes structures and error handling."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "shredword")), reason="the reference tree is not present on this machine")

SYMBOLS = ["create_trainer", "bpe_trainer_destroy", "bpe_load_corpus", "bpe_init", "bpe_count_bigrams", "bpe_merge_batch", "bpe_train", "bpe_save",
           "trainerCreate", "trainerDestroy", "addTextToTrainer", "preprocessTexts", "extractInitialSubwords", "computeLoss", "computeTokenLoss",
           "pruneVocabStep", "updateTokenScores", "trainUnigram", "getVocab", "saveVocab", "loadVocab"]


def _view(tmp_path, library):
    """<tmp>/shredword/{__init__,cbase,trainer}.py -> the reference's files, <tmp>/shredword/lib/libtrainer.so -> our library"""
    pkg = tmp_path / "shredword"
    (pkg / "lib").mkdir(parents=True)
    for name in os.listdir(os.path.join(REF, "shredword")):
        if name.endswith(".py"):
            os.symlink(os.path.join(REF, "shredword", name), pkg / name)
    os.symlink(library, pkg / "lib" / "libtrainer.so")
    return str(tmp_path)


def test_reference_package_binds_all_21_symbols_of_the_cuda_library(native, tmp_path):
    root = _view(tmp_path, native["lib"])
    code = ("import shredword, shredword.cbase as c, os; assert os.path.realpath(c._lib_path) == os.path.realpath(%r), c._lib_path; "
            "print(','.join(s for s in %r if getattr(c.lib, s, None) is not None)); print(shredword.__file__)") % (native["lib"], SYMBOLS)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=root), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[0].split(",") == SYMBOLS
    assert os.path.realpath(lines[1]).startswith(REF)  # it really was the reference's package


def test_reference_pytest_file_runs_unmodified_against_the_product_abi(native, tmp_path):
    from hostsim_lib import HS_SO
    root = _view(tmp_path, HS_SO)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", os.path.join(REF, "test", "test_bpe.py")],
                       env=dict(os.environ, PYTHONPATH=root, SHRED_QUIET="1"), capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0 and "3 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
