"""Debug tool: per-tile SM-clock timeline of one attention CTA (softmax groups A/B and the MMA thread).

Builds a -DSOME_ATTN_TRACE variant of the attention kernel into tools/_trace/ (git-ignored), runs one launch on
synthetic qkv (B clips x T frames) and prints, per key tile, when each role passed its barriers.  Not part of the product.
  python tools/attn_trace.py build      # here (no GPU)
  python tools/attn_trace.py run        # on the GPU box
"""
import ctypes, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.environ.get('ATTN_TRACE_OUT', os.path.join(HERE, '_trace', 'libattn_trace.so'))


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = os.path.join(ROOT, 'some_b200', 'csrc')
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
                           '-DSOME_ATTN_TRACE', *os.environ.get('ATTN_TRACE_FLAGS', '').split(), '-Xcompiler', '-fPIC', '-shared', '-o', OUT,
                           os.environ.get('ATTN_TRACE_SRC', os.path.join(src, 'attention_tc.cu')), os.path.join(src, 'host_common.cu'), '-lcudart'])


def run():
    import torch
    sys.path.insert(0, ROOT)
    from some_b200 import _lib
    lib = ctypes.CDLL(OUT)
    B, T = 64, 2584
    M = B * T
    torch.manual_seed(0)
    qkv = [torch.randn(M, 1536, device='cuda', dtype=torch.bfloat16) for _ in range(2)]
    out = [torch.empty(M, 512, device='cuda', dtype=torch.bfloat16) for _ in range(2)]
    cu = torch.arange(0, (B + 1) * T, T, device='cuda', dtype=torch.int32)
    trace = torch.zeros(4 * 64 * 4, device='cuda', dtype=torch.int64)
    a = _lib.AttnArgs()
    for g in range(2):
        a.qkv[g] = qkv[g].data_ptr()
        a.out[g] = out[g].data_ptr()
    a.groups, a.B, a.M, a.cu_frames, a.max_frames = 2, B, M, cu.data_ptr(), T
    lib.some_attention_set_trace.argtypes = [ctypes.c_void_p]
    lib.some_attention_varlen.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.some_attention_set_trace(trace.data_ptr()) == 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        assert lib.some_attention_varlen(ctypes.byref(a), st) == 0
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(4, 64, 4)
    t0 = t[t > 0].min()
    n = (T + 63) // 64
    print('tile | grp: s_full  max_done exp_done arrived | mma: before_p_full_wait  p_full_seen  pv_issued  qk_issued+commits   (SM clocks)')
    for j in range(min(n, 20)):
        g = j & 1
        e = t[g, j] - t0
        m = t[2, j] - t0
        print(f'{j:4d} |  {"AB"[g]}: {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d}')
    if t[3].any():   # v8 builds: role 3 = extra points
        print('tile | group started waiting for S_j | producer: waits for the slot of V_j, issues the load | mma thread: V_j seen | mma p_full_seen(j)')
        for j in range(4, min(n, 30)):
            x = t[3, j] - t0
            print(f'{j:4d} | {x[0]:7d} | {x[3]:7d} {x[2]:7d} | {x[1]:7d}  (issue -> seen {x[1] - x[2]:6d}) | {t[2, j, 1] - t0:7d}')
    for g in range(2):
        js = np.arange(g + 4, min(n, 40), 2)
        per = np.diff(t[g, js, 0]).mean()
        wait = (t[g, js[1:], 0] - t[g, js[:-1], 3]).mean()
        mx = (t[g, js, 1] - t[g, js, 0]).mean()
        ex = (t[g, js, 2] - t[g, js, 1]).mean()
        st = (t[g, js, 3] - t[g, js, 2]).mean()
        print(f'group {"AB"[g]}: period {per:.0f} clk/tile, wait-for-S {wait:.0f}, max pass {mx:.0f}, exp pass {ex:.0f}, store+arrive {st:.0f}')
    js = np.arange(4, min(n - 2, 38))
    m = t[2]
    print(f'mma thread per tile: idle in p_full wait {np.mean(m[js, 1] - m[js, 0]):.0f}, PV issue (4 MMAs) {np.mean(m[js, 2] - m[js, 1]):.0f}, '
          f'QK issue + 2 commits {np.mean(m[js, 3] - m[js, 2]):.0f}, kv_full wait etc. until next p_full wait {np.mean(m[js[1:], 0] - m[js[:-1], 3]):.0f}, '
          f'period {np.mean(np.diff(m[js, 0])):.0f}')
    for g in range(2):
        js = np.arange(g + 4, min(n - 2, 38), 2)
        print(f'group {"AB"[g]}: first-warp arrive -> mma saw p_full {np.mean(m[js, 1] - t[g, js, 3]):.0f} clk; '
              f'qk(j+2) issued+committed -> s_full(j+2) seen by the group {np.mean(t[g, js + 2, 0] - m[js, 3]):.0f} clk')


if __name__ == '__main__':
    {'build': build, 'run': run}[sys.argv[1]]()
