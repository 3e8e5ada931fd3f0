# usage (in the build container, then gpurun): bash scripts/build_k4_experiments.sh && gpurun -- 'bash scripts/gpu_k4_experiments.sh'
# Builds the default-off K4 experiments of csrc/k_pockets.cuh (DESIGN.md par. 9) as library variants under build/variants/ (they travel to the GPU box).
set -e
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false --default-stream per-thread -Xcompiler -fPIC -shared -Xcompiler -pthread -w"
mkdir -p build/variants
cp fpocket_b200/libfpocket_cuda.so build/variants/base.so
/usr/local/cuda/bin/nvcc $F -DFPK_K4_STAGE_TOUCHING=1 -o build/variants/stage.so fpocket_b200/csrc/capi.cu &
/usr/local/cuda/bin/nvcc $F -DFPK_K4_SEARCH_NEAR=1 -o build/variants/near.so fpocket_b200/csrc/capi.cu &
/usr/local/cuda/bin/nvcc $F -DFPK_K4_STAGE_TOUCHING=1 -DFPK_K4_SEARCH_NEAR=1 -o build/variants/stage_near.so fpocket_b200/csrc/capi.cu &
/usr/local/cuda/bin/nvcc $F -DFPK_K4_PROFILE -o build/variants/k4prof.so fpocket_b200/csrc/capi.cu &
wait
ls -la build/variants
