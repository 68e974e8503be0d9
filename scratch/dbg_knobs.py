import sys, os, subprocess
for mode in (1, 0):
  for d in (0, 4, 5, 1):
    env = dict(os.environ, SRK_TC_DBG=str(d), SRK_TC_MODE=str(mode), NIMG='64', ONLY='fprop')
    r = subprocess.run([sys.executable, 'scratch/prof_conv.py'], env=env, capture_output=True, text=True)
    print("mode=%d dbg=%d" % (mode, d), r.stdout.strip()[-200:], r.stderr.strip()[-300:], flush=True)
