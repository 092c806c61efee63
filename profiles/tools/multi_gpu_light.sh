# bash profiles/tools/multi_gpu_light.sh N   -- D2H ceiling, config 4 export and the 8K host-band frame on N GPUs
N=$1
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"; fi
$TR profiles/d2h_sweep.py > gpurun_out/r02_d2h_sweep_n$N.jsonl 2> gpurun_out/r02_d2h_sweep_n$N.err
$TR profiles/turntable_export.py > gpurun_out/r02_export_n$N.json 2> gpurun_out/r02_export_n$N.err
$TR profiles/band_host_8k.py > gpurun_out/r02_band_host_8k_n$N.json 2> gpurun_out/r02_band_host_8k_n$N.err
$TR profiles/band_8k.py > gpurun_out/r02_band_8k_n$N.json 2> gpurun_out/r02_band_8k_n$N.err
tail -n 1 gpurun_out/r02_*_n$N.json gpurun_out/r02_d2h_sweep_n$N.jsonl | cut -c1-600
