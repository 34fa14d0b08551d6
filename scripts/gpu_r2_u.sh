#!/bin/bash
mkdir -p gpurun_out
python scripts/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/u_parity.err
tail -40 gpurun_out/r02_parity_report.txt; tail -3 gpurun_out/u_parity.err
