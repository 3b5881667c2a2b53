import os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
for ns in (0, 1, 3, 5, 6):
    env = dict(os.environ, LRNDE_PROFILE_NSRC=str(ns))
    out = subprocess.run([sys.executable, os.path.join(here, "prof_feval.py"), "8192", "50"], env=env, capture_output=True, text=True)
    print("nsrc", ns, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
