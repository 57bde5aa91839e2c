# round 2, e2e slice-size A/B (one process)
timeout 200 python tests/e2e_ab.py 2>&1 | grep "sl0="
