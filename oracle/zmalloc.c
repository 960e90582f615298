/* oracle/zmalloc.c -- TEST INFRASTRUCTURE (not product code).
 *
 * Zero-fill malloc shim used to PIN the reference trainer's behaviour.
 * The reference allocates its linked-list nodes with malloc and never
 * initialises Symbol.deleted (reference shredword/csrc/bpe/histogram.cpp:14-19)
 * although it reads the flag in every scan (bpe.cpp:60,201,269).  Its output
 * therefore depends on stale heap bytes.  Running the unmodified reference with
 *     LD_PRELOAD=oracle/_ref/zmalloc.so
 * makes every malloc return zeroed memory, which is the only defined-behaviour
 * reading of the algorithm; this is the oracle all parity claims refer to
 * (SURVEY.md section 0.5 and Appendix B).
 */
#include <stddef.h>
extern void *__libc_calloc(size_t, size_t);
void *malloc(size_t n) { return __libc_calloc(1, n); }
