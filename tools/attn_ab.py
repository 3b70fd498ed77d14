"""Developer tool: interleaved A/B timing of attention-kernel variants (libraries built by tools/ab_bench.py / __graft_entry__).

  python tools/attn_ab.py base v9 v9r ...     # on the GPU box; 'base' = some_b200/libsome_b200.so, NAME = tools/_trace/lib_NAME.so
Every variant runs the C2 attention shape (64 clips x 2584 frames, two streams, 8 heads) ROUNDS times, interleaved with the others
so that box-to-box and clock drift cancel; prints mean / min ms per launch and checks the outputs against the first variant."""
import ctypes, os, sys
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from some_b200 import _lib  # noqa: E402

ROUNDS = int(os.environ.get('ATTN_AB_ROUNDS', '12'))


def main(names):
    libs = {}
    for n in names:
        path = os.path.join(ROOT, 'some_b200', 'libsome_b200.so') if n == 'base' else os.path.join(HERE, '_trace', f'lib_{n}.so')
        lib = ctypes.CDLL(path)
        lib.some_attention_varlen.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        libs[n] = lib
    B, T = 64, 2584
    M = B * T
    torch.manual_seed(0)
    qkv = [(torch.randn(M, 1536, device='cuda') * 1.5).to(torch.bfloat16) for _ in range(2)]
    cu = torch.arange(0, (B + 1) * T, T, device='cuda', dtype=torch.int32)
    outs = {n: [torch.empty(M, 512, device='cuda', dtype=torch.bfloat16) for _ in range(2)] for n in names}
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)

    def args(n):
        a = _lib.AttnArgs()
        for g in range(2):
            a.qkv[g] = qkv[g].data_ptr()
            a.out[g] = outs[n][g].data_ptr()
        a.groups, a.B, a.M, a.cu_frames, a.max_frames = 2, B, M, cu.data_ptr(), T
        return a

    a = {n: args(n) for n in names}
    times = {n: [] for n in names}
    for r in range(ROUNDS + 2):
        for n in names:
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = libs[n].some_attention_varlen(ctypes.byref(a[n]), st)
            e1.record()
            assert rc == 0, n
            torch.cuda.synchronize()
            if r >= 2:
                times[n].append(e0.elapsed_time(e1))
    ref = outs[names[0]]
    for n in names:
        t = torch.tensor(times[n])
        err = max((outs[n][g].float() - ref[g].float()).abs().max().item() for g in range(2))
        print(f'{n:10s} mean {t.mean():7.4f} ms  min {t.min():7.4f}  max {t.max():7.4f}  ( x4 launches = {4 * t.mean():6.3f} ms / step )  max|out - {names[0]}| = {err:.3e}')


if __name__ == '__main__':
    main(sys.argv[1:] or ['base'])
