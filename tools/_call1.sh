mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest.log 2>&1; tail -5 gpurun_out/gputest.log
python bench.py > gpurun_out/bench_ml20m_n1.json 2>gpurun_out/bench_ml20m_n1.err; tail -c 400 gpurun_out/bench_ml20m_n1.err
python tools/timeline.py step > gpurun_out/tl_step.txt 2>&1
python tools/timeline.py a > gpurun_out/tl_a.txt 2>&1; python tools/timeline.py d > gpurun_out/tl_d.txt 2>&1; python tools/timeline.py g > gpurun_out/tl_g.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
