run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/g_$name.json 2> gpurun_out/g_$name.err; }
run e1 NNSDP_EDGE_GROUP=1
run e2 NNSDP_EDGE_GROUP=2
run e4 NNSDP_EDGE_GROUP=4
run e8 NNSDP_EDGE_GROUP=8
