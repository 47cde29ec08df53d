mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "radon or c7g_active or kinetic_reactions or refill_kernels or c6_ion or test_specialized_kernel or flow or c4s_elm" > gpurun_out/r02_sanitizer.txt 2>&1
echo "sanitizer rc $?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_sanitizer.txt
tail -8 gpurun_out/r02_sanitizer.txt
