import os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
for extra in ({}, {"LRNDE_NO_CLUSTER": "1"}):
    for ns in (0, 1, 3, 5, 6):
        env = dict(os.environ, LRNDE_PROFILE_NSRC=str(ns), LRNDE_PROFILE_LAYERS="1", **extra)
        out = subprocess.run([sys.executable, os.path.join(here, "prof_feval.py"), "8192", "50"], env=env, capture_output=True, text=True)
        print(extra, "layer1 nsrc", ns, out.stdout.strip().split("TFLOP")[0] if out.stdout.strip() else out.stderr[-300:])
