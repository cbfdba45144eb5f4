python scripts/phase_clocks.py 4096 16 --warp-ends 2>&1 | head -14
