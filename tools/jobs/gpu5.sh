mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t5.txt 2>&1
tail -15 gpurun_out/r02_t5.txt
