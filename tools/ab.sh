for v in 8_8_4 4_12_4 8_12_4 4_16_4 6_10_4; do echo "== $v"; for m in 32 512; do BPE_CUDA_LIB=$PWD/build_ab/lib_$v.so timeout -s KILL 100 python tools/profile_run.py --merges $m --repeat 2 --profile-replace 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['train']; print(d['n_merges'], round(d['replace_ms'],2), 'ms', round(d.get('replace_gbs',0)), 'GB/s')"; done; BPE_CUDA_LIB=$PWD/build_ab/lib_$v.so timeout -s KILL 100 python tools/profile_run.py --kind 2 --merges 1 --repeat 3 --profile-replace 2>&1 | grep "run 2" | cut -c1-70; done
