timeout 600 python -m pytest tests/test_superpoint_gpu.py -m gpu -q 2>&1 | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/conv_halo_ab.py 2>&1 | tail -12
timeout 200 python tools/b1_profile.py 2>&1 | grep -E "opb profile|B=1" | tail -32
