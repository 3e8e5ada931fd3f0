for cfg in "512 4" "592 4" "888 4" "1184 4" "592 5"; do
  set -- $cfg
  FPK_CHUNK=$1 FPK_HOST_THREADS=$2 python bench.py --structures 10000 --steps 2 --warmup 1 --no-cpu --md-frames 0 --dpocket 0 2>gpurun_out/sweep_$1_$2.err | python -c "import json,sys; d=json.load(sys.stdin); print('chunk $1 threads $2 value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
