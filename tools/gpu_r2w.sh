mkdir -p gpurun_out
timeout 300 python tests/config3_timeline.py > gpurun_out/r2w_config3_timeline.txt 2>&1; echo "timeline rc=$?"
tail -22 gpurun_out/r2w_config3_timeline.txt
