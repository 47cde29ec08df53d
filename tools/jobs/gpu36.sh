mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_specialized_skeletons or test_specialized_kinetic_reactions or test_os_block_vector or test_c1_calcite" > gpurun_out/r02_racecheck.txt 2>&1
echo "racecheck rc $?"
grep -c "Race reported\|hazard" gpurun_out/r02_racecheck.txt
tail -6 gpurun_out/r02_racecheck.txt
