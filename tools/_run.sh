python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -2
echo default; for n in 1024 2048 8192; do N=$n python tools/chamfer_time.py; done
for v in vB vC vD vE vF vG vH; do echo $v; PCL_LIB_OVERRIDE=pointcloud_b200/_build/variants/$v.so python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -1; for n in 1024 2048 8192; do PCL_LIB_OVERRIDE=pointcloud_b200/_build/variants/$v.so N=$n python tools/chamfer_time.py; done; done
