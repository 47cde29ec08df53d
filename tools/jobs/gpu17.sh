mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t18.txt 2>&1
tail -8 gpurun_out/r02_t18.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -3 gpurun_out/r02_smoke.txt
