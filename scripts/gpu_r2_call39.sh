timeout 120 python scripts/probe_mn.py 2>&1 | tail -3
