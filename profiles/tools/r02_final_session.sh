python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1; tail -3 gpurun_out/r02k_pytest.log
ncu --set full --clock-control none --import-source on -k regex:k_tail -s 10 -c 1 -o gpurun_out/r02k_tail_suzy -f python profiles/perf_scenes.py suzy_suzy > gpurun_out/r02k_tail_prof.log 2>&1; tail -2 gpurun_out/r02k_tail_prof.log
ncu --set full --clock-control none --import-source on -k regex:"k_tri|k_super|k_resolve|k_xform" -s 8 -c 5 -o gpurun_out/r02k_frame -f python profiles/prof_geom.py 708 3840 2160 4 > gpurun_out/r02k_frame_prof.log 2>&1; tail -2 gpurun_out/r02k_frame_prof.log
python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; cat gpurun_out/r02k_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_under_ncu.log 2>&1
python profiles/perf_scenes.py > gpurun_out/r02k_perf_scenes.jsonl 2>&1; cat gpurun_out/r02k_perf_scenes.jsonl | cut -c1-330
