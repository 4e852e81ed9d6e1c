"""TMEM read / write throughput on one B200 SM (nf_selftest_tmem): bytes per clock of tcgen05.ld
(.x32 / .x16) and tcgen05.st with 4 / 8 / 16 warps, alone and under a tcgen05.mma stream, and the
MMA stream's slow-down under the reads.   python tools/tmem_bw.py  (gpurun) -> one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib  # noqa: E402


def main():
    ctx = _lib.default_context()
    it, mit = 2000, 2000
    out = {}
    _lib.selftest_tmem(ctx, 4, 10, 10, 5)
    for w in (4, 8, 16):
        for name, mode in (('ld32', 1), ('ld16', 9), ('st16', 2)):
            c, _ = _lib.selftest_tmem(ctx, w, it, 0, mode)
            out['%s_w%d' % (name, w)] = {'cycles': c, 'B_per_clk': w * it * 8192 / max(c, 1)}
    _, c = _lib.selftest_tmem(ctx, 0, 0, mit, 4)
    out['mma_n128_alone'] = {'cycles': c, 'clk_per_mma': c / (mit * 8)}
    _, c = _lib.selftest_tmem(ctx, 0, 0, mit, 4 | 16)
    out['mma_n256_alone'] = {'cycles': c, 'clk_per_mma': c / (mit * 8)}
    for w in (4, 8):
        for name, mode in (('ld32', 1), ('st16', 2), ('ld32_st16', 3)):
            cr, cm = _lib.selftest_tmem(ctx, w, it, mit, mode | 4)
            nbytes = w * it * 8192 * (2 if mode == 3 else 1)
            out['%s_w%d_with_mma' % (name, w)] = {
                'reader_cycles': cr, 'B_per_clk': nbytes / max(cr, 1), 'mma_cycles': cm,
                'clk_per_mma': cm / (mit * 8)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
