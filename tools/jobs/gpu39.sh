mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t39.txt 2>&1
tail -5 gpurun_out/r02_t39.txt
