#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s12_pytest.txt
python tools/run_config.py C5 --sample 20 --reps 2 > gpurun_out/s12_c5.txt 2>&1
python tools/run_config.py C3 --sample 100 --reps 3 > gpurun_out/s12_c3.txt 2>&1
CARETTA_B200_STREAMS=1 python tools/run_config.py C5 --sample 0 --reps 1 > gpurun_out/s12_c5_serial.txt 2>&1
tail -4 gpurun_out/s12_pytest.txt; cat gpurun_out/s12_c5.txt gpurun_out/s12_c3.txt gpurun_out/s12_c5_serial.txt
