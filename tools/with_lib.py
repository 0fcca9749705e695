"""Run a tool against another build of libspq_b200.so (same-box A/B of kernel changes):
    SPQ_LIB=/path/to/old/libspq_b200.so python tools/with_lib.py tools/gemm_bench.py 32768 2304 768 50 f32
Box-to-box variance on the pool is ~3-5 %, so two builds are only comparable inside one gpurun call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_qat_on_gpt2_b200 import _lib

if os.environ.get("SPQ_LIB"):
    _lib.load_library(os.environ["SPQ_LIB"])
sys.argv = sys.argv[1:]
exec(compile(open(sys.argv[0]).read(), sys.argv[0], "exec"))
