# BASELINE configs[2] on N GPUs of one box (bench.py's clip leg: 512 x 3x768x768, frames sharded, NCCL all-gather of the outputs)
# usage: gpurun --gpus N -- 'bash tests/run_clip.sh N'
N=${1:-2}
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-hbm-leg > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
echo "rc=$?"; grep -c "NCCL INFO" gpurun_out/r2_scale_n$N.err; grep -m3 "NVLS\|Connected all rings\|nranks" gpurun_out/r2_scale_n$N.err | cut -c1-200
python -c "
import json; b=json.loads(open('gpurun_out/r2_scale_n$N.json').read().strip().splitlines()[-1]); print(b['n_gpus'], b['value'], b['ms_per_step']); print(json.dumps(b['clip']))"
