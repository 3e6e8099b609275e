bash profiles/capture.sh r2_final_harvest5_b4096 ssd_kernel 30 harvest5 4096
bash profiles/capture.sh r2_final_harvest5_b65536 ssd_kernel 30 harvest5 65536
bash profiles/capture.sh r2_final_cleanup5_b4096 ssd_kernel 30 cleanup5 4096
bash profiles/capture.sh r2_final_cleanup10_b16384 ssd_kernel 30 cleanup10 16384
bash profiles/capture.sh r2_final_reset_harvest5_b4096 ssd_kernel 20 harvest5 4096 reset
bash profiles/capture.sh r2_final_render_harvest5_b4096 ssd_kernel 20 harvest5 4096 render
bash profiles/capture.sh r2_final_frontend7_b4096 obs_frontend 4 frontend7 4096
python bench.py --steps 20 --warmup 5 --no-extra --no-train --no-cpu-baseline > gpurun_out/r2_launchlist_bench.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 20 --warmup 5 --no-extra --no-train --no-cpu-baseline > gpurun_out/r2_launchlist_ncu.log 2>&1
rm -f gpurun_out/*.ncu-rep
ls gpurun_out | tail -30
