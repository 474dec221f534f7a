mkdir -p gpurun_out
for d in 0 3 4 7 2048 2176 2192 2208 2052 2180 131 135 128; do
  echo "== dbg=$d"
  env LFDM_CONV_DBG=$d timeout 100 python tools/profile_kernels.py --only c --iters 7 2>>gpurun_out/mb.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['kernel'] in ('conv3x3_c64_32_nogn', 'conv3x3_c128_16', 'conv3x3_c256_8', 'qkv_c64_32', 'out_c256_32', 'conv3x3_c512_4'): print('   ', d['kernel'], d['ms'])
"
done
