run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload mid-W100-D50-beta2-Q1024 --steps 5 --warmup 3 --no-e2e --no-extras > gpurun_out/m_$name.json 2> gpurun_out/m_$name.err; }
run base X=1
run pn NNSDP_PANEL_NARROW=1
run pn4 NNSDP_PANEL_NARROW=1 NNSDP_PANEL_GROUP=4
run pn8 NNSDP_PANEL_NARROW=1 NNSDP_PANEL_GROUP=8
NNSDP_PANEL_NARROW=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "W100 or wide or programs or config2" 2>&1 | tail -2
