import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import train_step_bench as T
import __graft_entry__ as g
g.build()
T.run(64, steps=2, warmup=1)
