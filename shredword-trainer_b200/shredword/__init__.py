"""B200-native drop-in for the BPE half of the `shredword` trainer package (reference shredword/__init__.py)."""
from .encoder import BPEEncoder
from .trainer import BPETrainer, UnigramTrainer

__version__ = "0.1.0+b200"
