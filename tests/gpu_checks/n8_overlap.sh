# N=8: default step (optimizer reads the bf16 payload) and the phased backward / overlapped all-reduce variants
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_optimizer.py -q -x -k payload 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3"
timeout 200 $TR > gpurun_out/r2_bench_n8_ftalign_v2.log 2>&1
tail -1 gpurun_out/r2_bench_n8_ftalign_v2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'], {k[:28]: round(v,3) for k,v in d.get('step_breakdown_ms',{}).items()})" || tail -5 gpurun_out/r2_bench_n8_ftalign_v2.log
for spec in "9,5 0" "6 0" "9,5 16"; do
  set -- $spec
  tag=$(echo $1 | tr ',' '_')_sm$2
  timeout 200 $TR --no_e2e --profile_steps 0 --overlap_cuts $1 --overlap_sms $2 > gpurun_out/r2_n8_overlap_$tag.log 2>&1
  tail -1 gpurun_out/r2_n8_overlap_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$spec', round(d['value']), round(d['ms_per_step'],3), round(d['loss'],5), {k[:28]: round(v,3) for k,v in d.get('step_breakdown_ms',{}).items()})" || tail -5 gpurun_out/r2_n8_overlap_$tag.log
done
