# bash profiles/tools/multi_gpu_session.sh N   -- the multi-GPU measurements of one box with N GPUs, one JSON(L) file each
# under gpurun_out/ (copied to profiles/ afterwards).  Every multi-rank command: one process per GPU via torchrun.
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
$TR --master-port 29601 profiles/d2h_sweep.py > gpurun_out/r02_d2h_sweep_n$N.jsonl 2> gpurun_out/r02_d2h_sweep_n$N.err
$TR --master-port 29602 profiles/turntable_export.py > gpurun_out/r02_export_n$N.json 2> gpurun_out/r02_export_n$N.err
$TR --master-port 29603 profiles/band_host_8k.py > gpurun_out/r02_band_host_8k_n$N.json 2> gpurun_out/r02_band_host_8k_n$N.err
$TR --master-port 29604 profiles/band_8k.py > gpurun_out/r02_band_8k_n$N.json 2> gpurun_out/r02_band_8k_n$N.err
$TR --master-port 29605 bench.py --gpus $N --steps 64 --warmup 8 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -n 2 gpurun_out/r02_*_n$N.json gpurun_out/r02_d2h_sweep_n$N.jsonl | cut -c1-700
