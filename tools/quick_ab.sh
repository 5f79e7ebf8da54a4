timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/pb_$name.json 2> gpurun_out/pb_$name.err; }
run b4 NNSDP_PANEL_BATCH=4
run b8 NNSDP_PANEL_BATCH=8
run b16 NNSDP_PANEL_BATCH=16
run b32 NNSDP_PANEL_BATCH=32
