timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 400 python tools/ab_test.py 10 2>&1 | grep -A1 "identity_diag"
timeout 200 python tools/b1_profile.py 2>&1 | grep -E "B=1 match|mlp3"
