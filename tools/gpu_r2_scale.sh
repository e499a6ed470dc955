#!/bin/bash
# bench.py under torchrun at N GPUs (+ the reference arm), for the record
N=${1:-8}
tag=${2:-r2h}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${tag}_ref_n$N.json 2> gpurun_out/${tag}_ref_n$N.err; echo "ref rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms %.2f solve %.2f gather_ms %s GB/s %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["ms_solve"], d.get("final_gather_ms"), d.get("final_gather_gbs_into_rank0")))
print("e2e", json.dumps(d["e2e"]))
for o in d.get("other_configs", []): print(o["workload"][:30], "%.4g" % o["value"], "e2e %.4g" % o["e2e"]["value"], o.get("final_gather_ms"), o["systems_ok"])
print("strong", [(s["total_systems"], "%.4g" % s["value"], s["final_gather_ms"]) for s in d.get("strong_scaling", [])])
r = json.loads(open("gpurun_out/${tag}_ref_n$N.json").read().strip().splitlines()[-1])
print("ref value %.4g cores %s" % (r["value"], r["cpu_baseline"]["cores"]))
PY
