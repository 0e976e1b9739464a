# A/B of two builds of the library on the SAME box: current tree vs profiles/micro/_ab/libpillars_base.so (interleaved)
# To produce the base build: git stash (or checkout the commit to compare against); python -m lidar_vision_vqa_b200.build; mkdir -p profiles/micro/_ab;
# cp lidar_vision_vqa_b200/_build/libpillars_b200.so profiles/micro/_ab/libpillars_base.so; git stash pop; rebuild.  (*.so files are git-ignored.)
L=lidar_vision_vqa_b200/_build/libpillars_b200.so
cp $L /tmp/new.so
one() { python bench.py --steps 20 --warmup 5 --no-cpu --no-extra-workloads --no-extractor 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']; print('$1 group_ms %.4f features_ms %.4f scatter_ms %.4f serial %.4f value %.0f' % (s['group_ms'], s['features_ms'], s['scatter_ms'], s['serial_ms_per_step'], d['value']))"; }
for i in 1 2 3; do
  cp /tmp/new.so $L; touch $L; one new
  cp profiles/micro/_ab/libpillars_base.so $L; touch $L; one base
done
