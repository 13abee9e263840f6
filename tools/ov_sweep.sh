export CIM_OVERLAP_NOCHECK=1
python tools/bench_overlap.py tiled 2>&1 | tail -1
for f in NOFENCE NOLOAD NOMMA NOEPI "NOEPI -DCIM_OV_ABL_NOFENCE -DCIM_OV_ABL_NOLOAD"; do
  touch cim_b200/csrc/mask_overlap_tc.cu; make -C cim_b200/csrc EXTRA="-DCIM_OV_ABL_$f" > /dev/null 2>&1 || echo build failed
  echo "ablation $f"; python tools/bench_overlap.py tiled 2>&1 | tail -1
done
