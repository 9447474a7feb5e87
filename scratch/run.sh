set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_relaxed.py -m gpu -x -q -k "roots_next_to_a_centre or flat_triangles" > gpurun_out/r3e_pytest_new.log 2>&1; tail -5 gpurun_out/r3e_pytest_new.log
GCS_B200_LIB=$PWD/build/alt_unstable_h2.so python -m pytest tests/test_gpu_relaxed.py -m gpu -q -k "roots_next_to_a_centre" > gpurun_out/r3e_pytest_unstable.log 2>&1; tail -12 gpurun_out/r3e_pytest_unstable.log
python scratch/kbench.py 5 1 524288 8 > gpurun_out/r3e_kbench.log 2>&1
python scratch/kbench.py 5 1,3 1048576 8 >> gpurun_out/r3e_kbench.log 2>&1
python scratch/kbench.py 8 1,3,5 524288 2 >> gpurun_out/r3e_kbench.log 2>&1
python scratch/kbench.py 5 1,5 524288 2 >> gpurun_out/r3e_kbench.log 2>&1
cat gpurun_out/r3e_kbench.log
