run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/w_$name.json 2> gpurun_out/w_$name.err; }
run wm4 NNSDP_IBP_WM=4
run wm2 NNSDP_IBP_WM=2
run wm1 NNSDP_IBP_WM=1
