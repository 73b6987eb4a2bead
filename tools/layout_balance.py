import time, os, sys
sys.path.insert(0, '/root/repo')
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
p = make_problem("venice", 1)
for pw in ("-1", "0", "2", "3", "5", "8"):
    if pw == "-1": os.environ["B200BA_LAYOUT_PW"] = "1000000"     # ~ equal point counts (the old split)
    else: os.environ["B200BA_LAYOUT_PW"] = pw
    ms = []
    for i in range(4):
        s = binding.layout_stats(p, 64); ms.append(s["ms"])
    print("pw", pw, "layout ms", [round(float(x), 1) for x in ms], "conflict", round(float(s["conflict_factor"]), 4))
