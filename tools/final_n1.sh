set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_final.log 2>&1; tail -3 gpurun_out/pytest_r2_final.log
python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
python tools/timeline.py final > gpurun_out/timeline_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-inference --no-stress > gpurun_out/b_r2_ncu.log 2>&1
bash tools/ncu_capture.sh r2 > gpurun_out/ncu_capture.log 2>&1

ls -la gpurun_out/ncu_r2_*.csv | head
