timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/v_quick.json 2> gpurun_out/v_quick.err
