# Round 2, later session: parity first, then same-box A/B of the previous build (alt/lib_base.so) against the new one,
# instruction counts of the frame kernels.   bash profiles/tools/r02_session_b.sh TAG
tag=${1:-r02m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
bash profiles/tools/ab3.sh rust-sloth_b200/alt/lib_base.so rust-sloth_b200/libsloth_b200.so rust-sloth_b200/alt/lib_base.so rust-sloth_b200/libsloth_b200.so 2>&1 | tee gpurun_out/${tag}_ab.log
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_tri|k_super|k_resolve|k_xform|k_tail" -s 8 -c 6 --csv --log-file gpurun_out/${tag}_inst.csv python profiles/prof_geom.py 708 3840 2160 4 > gpurun_out/${tag}_inst.log 2>&1; cut -d, -f5,13- gpurun_out/${tag}_inst.csv | tail -20
