"""One optimize_parameters step of a golden case, for compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_step.py c1_affine64 bf16 auto
    compute-sanitizer --tool memcheck  python scripts/sanitize_step.py c1_affine64 fp32 generic
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import helpers as H  # noqa: E402

name, precision, engine = (sys.argv[1:4] + ["c1_affine64", "bf16", "auto"][len(sys.argv) - 1:])[:3]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
model, cfg, states, (A, B) = H.build_case(name, precision=precision, conv_engine=engine)
print("losses", H.run_engine_steps(model, A, B, steps)[-1])
print("SANITIZE_STEP_DONE", name, precision, engine)
