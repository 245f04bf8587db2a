#!/bin/bash
# A/B of the host-expanded end-to-end leg on N GPUs of one box: default, spinning copy waits, the old two-slot / 256-line tasks
N=${1:-4}
mkdir -p gpurun_out
run() { env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-mismatch --no-cli --e2e-steps 3 2>gpurun_out/x.err >gpurun_out/x.json
python -c "
import json; l=json.load(open('gpurun_out/x.json')); print('$*', 'e2e host', round(l['e2e_host_expand']['ms_per_step'],1), 'dev', round(l['e2e_device_formatter']['ms_per_step'],1), 'packed', round(l['e2e_packed']['ms_per_step'],1))"; }
run PGS_X=1
run PGS_SPIN_SYNC=1
run PGS_HX_RING=1 PGS_HX_TASK_LINES=256 PGS_SPIN_SYNC=1
run PGS_X=1
