#!/bin/bash
# usage: tools/scale_sweep.sh N  -- the BASELINE.json sequence configs 4 and 5 on N GPUs of this box (strong scaling: fixed clips)
N=$1
run() {  # name, extra args
  local out=gpurun_out/r2_scale_$1_n$N
  shift_args=("${@:2}")
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "${shift_args[@]}" > $out.json 2> $out.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "${shift_args[@]}" > $out.json 2> $out.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("$out.json"))
    print("$1 N=$N value %.2f e2e %.2f ms/step %.1f checksums %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["checksums"]["digest_first_block"] + ("" if d["checksums"]["identical_across_ranks"] else " MISMATCH")))
except Exception as e:
    print("$1 N=$N FAILED", e)
PY
}
run c4_inpaint_ns_4k_300f --workload inpaint_ns_4k --clip-frames 300 --steps 3 --warmup 3 --no-cpu
run c5_farneback_8k_l5_1000f --workload farneback_8k_l5 --clip-frames 1001 --steps 2 --warmup 3 --no-cpu --no-parity
