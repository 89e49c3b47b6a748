REPS=1 timeout 120 python tools/proj_only.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project -c 1 -f -o gpurun_out/r2v_k_project python tools/proj_only.py > gpurun_out/r2v_ncu_proj.log 2>&1
ncu -i gpurun_out/r2v_k_project.ncu-rep --page raw --csv > gpurun_out/r2v_proj_raw.csv 2>/dev/null
