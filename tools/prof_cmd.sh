set -e
cd /root/repo
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/plain_k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcb_env_kernel -s 10 -c 1 -o gpurun_out/prof_r01k -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-her > gpurun_out/ncu_k.log 2>&1
ncu -i gpurun_out/prof_r01k.ncu-rep --page raw --csv > gpurun_out/prof_raw_k.csv
ncu -i gpurun_out/prof_r01k.ncu-rep --page source --csv > gpurun_out/prof_src_k.csv
tail -2 gpurun_out/ncu_k.log
