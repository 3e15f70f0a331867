"""BASELINE.json configs[2] and configs[3] at full size on one B200 (secondary bench lines):

    python scripts/run_configs.py kmeans     # = python bench.py --config cfg3
    python scripts/run_configs.py filtered   # = python bench.py --config cfg4
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, cfg in (("kmeans", "cfg3"), ("filtered", "cfg4")):
    if which in (name, "all"):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "bench.py"), "--config", cfg, "--steps", "10"])
