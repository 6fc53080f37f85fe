import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from golden_util import load_golden, steps_of
import qd_oracle as O
from test_gpu_archives import make_gpu_archive
from test_oracle_golden import make_oracle_archive
case = load_golden("add_traj")["single_mae"]
g, o = make_gpu_archive(case["cfg"]), make_oracle_archive(case["cfg"])
b = steps_of(case)[0]["in"]
for i in range(len(b["objective"])):
    row = {k: v[i] for k, v in b.items()}
    ig, io = g.add_single(**row), o.add_single(**row)
    be_g, be_o = g.best_elite, o.best_elite
    ok = (be_g is None and be_o is None) or np.array_equal(np.asarray(be_g["solution"]), np.asarray(be_o["solution"]))
    if not ok or ig["status"] != io["status"]:
        last = g._store._last
        print("diverged at", i, "status", ig["status"], io["status"], "obj", row["objective"], "gpu best", be_g["objective"], be_g["index"], "ora best", be_o["objective"], be_o.get("index"),
              "res.best_cell", last.best_cell, "n_winners", last.n_winners, "flags", last.flags, "best_objective", last.best_objective, "obj_max", last.obj_max)
        break
else:
    print("no divergence over", len(b["objective"]), "adds")
