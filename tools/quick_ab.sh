run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/pc_$name.json 2> gpurun_out/pc_$name.err; }
run b2 NNSDP_PANEL_BATCH=2
run b8 NNSDP_PANEL_BATCH=8
run g4b8 NNSDP_PANEL_BATCH=8 NNSDP_PANEL_GROUP=4
NNSDP_PANEL_BATCH=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stress or ragged or wide_nets or panel" 2>&1 | tail -2
