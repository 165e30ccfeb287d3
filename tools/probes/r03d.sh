# r03d: longer GPU differential fuzz with fresh seeds (CUDA path vs oracle, then vs the transliterated reference)
mkdir -p gpurun_out
python tools/fuzz_kernel_vs_oracle.py --backend gpu --master-seed 1001 4 > gpurun_out/r03d_fuzz_gpu_vs_oracle.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r03d_fuzz_gpu_vs_oracle.log
python tools/fuzz_kernel_vs_oracle.py --backend gpu --engine ref --master-seed 1002 3 > gpurun_out/r03d_fuzz_gpu_vs_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r03d_fuzz_gpu_vs_ref.log
