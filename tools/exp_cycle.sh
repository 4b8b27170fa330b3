cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
bash tools/exp_bench.sh pick 2>&1 | tail -4
