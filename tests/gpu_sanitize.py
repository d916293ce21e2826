"""Driver for compute-sanitizer: one forward + backward of a golden scene through every tensor-core kernel of the training path.
Usage: compute-sanitizer --tool memcheck python tests/gpu_sanitize.py [scene] [train:0|1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
from test_gpu_backward import run_backward  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tennis_dense"
training = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
os.environ["PE_BWD_TC"] = "1"
golden, loss, got_in, got_par = run_backward(name, training, precision="mixed")
print("done", name, training, loss)
