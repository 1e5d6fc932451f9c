import sys, os
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
lgr = ge._load_package()
ex = lgr.make_executor(64, 256)
for w, name in ((1, "montgomery"), (5, "shoup"), (0, "imad_wide"), (3, "imad_lo")):
    print(name, "%.4g" % ex.ubench(w))
