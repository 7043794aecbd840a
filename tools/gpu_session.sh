set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:cost_volume_topk -c 1 -f -o gpurun_out/prof_costvol2 python tools/ncu_target.py 1 hot > gpurun_out/ncu_cv2.log 2>&1; tail -3 gpurun_out/ncu_cv2.log
