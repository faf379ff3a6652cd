set -u
mkdir -p gpurun_out
python tools/gemm_one.py 4416 3072 768 256 1 0 0 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/gemm_ffn1 python tools/gemm_one.py 4416 3072 768 256 1 0 0 > gpurun_out/ncu_gemm_ffn1.log 2>&1
ncu -i gpurun_out/gemm_ffn1.ncu-rep --page source --csv > gpurun_out/gemm_ffn1_src.csv 2>/dev/null
ncu -i gpurun_out/gemm_ffn1.ncu-rep --page raw --csv > gpurun_out/gemm_ffn1_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/gemm_o python tools/gemm_one.py 4416 768 768 192p 0 1 1 > gpurun_out/ncu_gemm_o.log 2>&1
ncu -i gpurun_out/gemm_o.ncu-rep --page source --csv > gpurun_out/gemm_o_src.csv 2>/dev/null
ncu -i gpurun_out/gemm_o.ncu-rep --page raw --csv > gpurun_out/gemm_o_raw.csv 2>/dev/null
ls -la gpurun_out/gemm_*; tail -3 gpurun_out/ncu_gemm_ffn1.log
