"""Three training steps at BASELINE C3 (1024 rays, 64+128) for `ncu` launch lists / captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/X_train_launches.csv python tools/train_one_step.py
(the last third of the launches = one warmed-up step)."""
import os
import sys

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import __graft_entry__ as g  # noqa: E402
import train_step_bench as T  # noqa: E402

g.build()
T.run(int(sys.argv[1]) if len(sys.argv) > 1 else 128, steps=2, warmup=1)
