set -x
mkdir -p gpurun_out
NMRF_B200_LIB=nmrf_b200/libnmrf_b200_trace.so timeout 300 python tools/ra_trace.py > gpurun_out/ra_trace.log 2>&1; tail -3 gpurun_out/ra_trace.log
