# Same-box A/B of the FP64 tensor-core paths: switches through the environment, alternate builds of the library through
# NNSDP_B200_LIB (build kernels_dgemm.cu with -DNNSDP_DGEMM_TN / _MINB / _WN / _LDA_PAD / _DK, kernels_gram.cu with
# -DNNSDP_GRAM_WARPS_N / _LD_PAD, link with the other objects of nn-sdp_b200/build into nn-sdp_b200/lib/alt_<name>.so).
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-extras > gpurun_out/dg_$name.json 2> gpurun_out/dg_$name.err
  env "$@" timeout 300 python tools/crown_timing.py 2>&1 | grep W1000 > gpurun_out/dg_$name.crown
}
run default X=1
run no_dmma_ibp NNSDP_NO_DMMA_IBP=1
run no_affine_layers NNSDP_NO_AFFINE_LAYERS=1
for lib in nn-sdp_b200/lib/alt_*.so; do
  [ -f "$lib" ] && run "$(basename $lib .so)" NNSDP_B200_LIB=/root/repo/$lib
done
