mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/c17_bench.json 2>gpurun_out/c17_bench.err; tail -c 300 gpurun_out/c17_bench.err; python tools/_show.py gpurun_out/c17_bench.json
timeout 300 python bench.py --impl reference > gpurun_out/c17_ref.json 2>gpurun_out/c17_ref.err; tail -c 300 gpurun_out/c17_ref.err; cut -c1-400 gpurun_out/c17_ref.json
