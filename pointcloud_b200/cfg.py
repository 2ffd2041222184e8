"""Loss hyper-parameters the hot path reads (mirror of pointcloud_vision/cfg.py:13-41; module-level
globals, imported as `cfg`, overridable by the caller exactly like the reference's)."""

device = 'cuda'
precision = '16-mixed'   # cfg.py:13 -- pred may arrive as fp16/bf16; the kernels up-cast like .float()
debug = False            # cfg.py:16 -- enables the unassigned-ratio check of utils.py:261-265

# Earth Mover's Distance loss precision (cfg.py:36-41)
emd_eps = 0.005
emd_iterations = 50
emd_test_eps = 0.002
emd_test_iterations = 10000

# Chamfer arithmetic: 'unfused' = pytorch3d CPU build order, 'fma' = nvcc-contracted knn.cu order
chamfer_mode = 'unfused'
