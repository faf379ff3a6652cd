# source-level ncu capture of the FFN1 (+GELU, LayerNorm folded) GEMM launch; run on the GPU box
set -u
mkdir -p gpurun_out
python tools/gemm_one.py 4416 3072 768 256 1 0 0 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_gemm_ffn1 python tools/gemm_one.py 4416 3072 768 256 1 0 0 1 > gpurun_out/r02_ncu_gemm_ffn1.log 2>&1
ncu -i gpurun_out/r02_gemm_ffn1.ncu-rep --page source --csv > gpurun_out/r02_gemm_ffn1_src.csv 2>/dev/null
ncu -i gpurun_out/r02_gemm_ffn1.ncu-rep --page raw --csv > gpurun_out/r02_gemm_ffn1_raw.csv 2>/dev/null
ls -la gpurun_out/r02_gemm_*; tail -3 gpurun_out/r02_ncu_gemm_ffn1.log
