python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err; tail -c 300 gpurun_out/r2i_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2i_default_cmd_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_default_cmd.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
