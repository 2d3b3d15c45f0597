timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r38_bench.json 2> gpurun_out/r38_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r38_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r38_bench_ref.json 2> gpurun_out/r38_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r38_bench_ref.json | cut -c1-400
