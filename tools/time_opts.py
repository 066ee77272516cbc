"""A/B timing of the headline kernel's formulations (library built with OFDM_FAST_EXPERIMENTS=1):
    python tools/time_opts.py [N] [order] [symbols] [reps]
Runs tools/time_fused.py once per OFDM_B200_FAST_OPT value in a fresh process (the selection is read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:] or ["1024", "64", "162761", "20"]
for opt, sync in (("0", None), ("1", None), ("2", None), ("3", None), ("3", "x")):
    env = dict(os.environ, OFDM_B200_FAST_OPT=opt, OFDM_B200_FAST_TAPS8="1")
    if sync:
        env["OFDM_B200_FAST_SYNC"] = "9"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_fused.py"), *args], env=env, capture_output=True, text=True)
    print(f"opt={opt} sync={'other' if sync else 'default'}: {out.stdout.strip()} {out.stderr.strip()[-300:]}", flush=True)
