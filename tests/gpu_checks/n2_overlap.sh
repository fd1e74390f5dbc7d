# N=2: optimizer reading the bf16 payload directly; phased backward / overlapped gradient all-reduce sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no_e2e --profile_steps 0"
for spec in "off 0 bf16" "off 0 f32" "9,5 0 bf16"; do
  set -- $spec
  tag=$(echo $1 | tr ',' '_')_sm$2_$3
  timeout 200 $TR --overlap_cuts $1 --overlap_sms $2 --grad_payload $3 > gpurun_out/r2_n2_overlap_$tag.log 2>&1
  tail -1 gpurun_out/r2_n2_overlap_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$spec', round(d['value']), round(d['ms_per_step'],3), round(d['loss'],5), {k[:28]: round(v,3) for k,v in d.get('step_breakdown_ms',{}).items()})" || tail -5 gpurun_out/r2_n2_overlap_$tag.log
done
