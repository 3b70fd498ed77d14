"""Developer A/B harness: run bench.py against kernel-variant builds of the library (tools/_trace/lib_<name>.so).

  python tools/ab_bench.py build NAME [--src FILE.cu] [--replace gemm] [-DFLAG ...]
        # here: rebuilds ONE object with the flags — attention_tc.cu by default, FILE.cu with --src (it must export the same
        # entry points), or csrc/<X>.cu with --replace X — and links it with the other csrc/build/*.o
  python tools/ab_bench.py run NAME [bench args]      # on the GPU box: bench.py with that library
  python tools/ab_bench.py pytest NAME [pytest args]  # on the GPU box: the test-suite against that library
The product never loads these: this script repoints some_b200._lib.LIB_PATH for its own process only."""
import glob, os, pathlib, runpy, subprocess, sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = ROOT / 'some_b200' / 'csrc'


def build(name, flags):
    out = HERE / '_trace'
    out.mkdir(exist_ok=True)
    obj = out / f'variant_{name}.o'
    src = CSRC / 'attention_tc.cu'
    replaced = 'attention_tc'
    if '--replace' in flags:
        i = flags.index('--replace')
        replaced = flags[i + 1]
        src = CSRC / f'{replaced}.cu'
        flags = flags[:i] + flags[i + 2:]
    if '--src' in flags:
        i = flags.index('--src')
        src = pathlib.Path(flags[i + 1]).resolve()
        flags = flags[:i] + flags[i + 2:]
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
                           '-Xcompiler', '-fPIC', *flags, '-c', str(src), '-o', str(obj)])
    objs = [o for o in glob.glob(str(CSRC / 'build' / '*.o')) if not o.endswith(f'/{replaced}.o')]
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', str(out / f'lib_{name}.so'),
                           str(obj), *objs, '-lcudart'])


def run(name, args):
    sys.path.insert(0, str(ROOT))
    from some_b200 import _lib
    _lib.LIB_PATH = HERE / '_trace' / f'lib_{name}.so'
    sys.argv = ['bench.py', *args]
    runpy.run_path(str(ROOT / 'bench.py'), run_name='__main__')


def run_pytest(name, args):
    sys.path.insert(0, str(ROOT))
    import pytest
    from some_b200 import _lib
    _lib.LIB_PATH = HERE / '_trace' / f'lib_{name}.so'
    os.chdir(ROOT)
    sys.exit(pytest.main(args or ['tests/test_gpu_kernels.py', 'tests/test_gpu_pipeline.py', '-m', 'gpu', '-x', '-q']))


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == 'pytest':
        run_pytest(sys.argv[2], sys.argv[3:])
    else:
        run(sys.argv[2], sys.argv[3:])
