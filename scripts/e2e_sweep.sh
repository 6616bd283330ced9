for mb in 256 128 64 32 16; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tune copy_chunk_mb=$mb 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk_mb', sys.argv[1], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), 'pcm16', round(d['e2e_pcm16']['value']), round(d['e2e_pcm16']['ms_per_step'],1))" $mb; done
