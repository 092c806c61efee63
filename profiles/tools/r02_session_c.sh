# Round 2, closing session of a build: parity (all GPU tests), bench line, launch list of the bench command, ncu --set full
# of one frame's kernels, the other configs.   bash profiles/tools/r02_session_c.sh TAG
tag=${1:-r02q}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-1500 gpurun_out/${tag}_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tri|k_super|k_resolve|k_xform" -s 8 -c 5 -o gpurun_out/${tag}_frame -f python profiles/prof_geom.py 708 3840 2160 4 > gpurun_out/${tag}_frame_prof.log 2>&1; tail -1 gpurun_out/${tag}_frame_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 600 python profiles/perf_scenes.py > gpurun_out/${tag}_perf_scenes.jsonl 2>&1; cut -c1-260 gpurun_out/${tag}_perf_scenes.jsonl
