#!/bin/bash
# Level-0 sweep times of the headline workload for alternative Heat1D team shapes (profiles/r01m_team_shapes.txt).
# The run that produced the profile had ('Heat1D', 128, 9) and ('Heat1D', 96, 11) added to SHAPES in
# pymgrit_b200/build.py and an MGB_TEAM_SHAPE=T,E override in core/device_level.py:team_shape; both were removed again
# (one warp per system is 2-3x faster), so re-running this needs those two edits.
mkdir -p gpurun_out
for shape in "32,33" "128,9" "96,11"; do
  echo "== shape $shape"
  MGB_TEAM_SHAPE=$shape python bench.py --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('ms_per_step %.3f  iterations %d  conv %s' % (d['ms_per_step'], d['config']['iterations'], d['config']['conv']))
        for k in d['kernels']:
            print('  %-45s %.3f ms  %.0f GB/s' % (k['name'], k['ms'], k['gbs']))
    elif 'rror' in line: print(line.strip()[:300])
"
done
