"""tests/golden/cfg_parse.json: a .cfg text exercising the key set of the reference's FPL+ configs
(config_dual/data_vs/*.cfg) and what the REFERENCE's own parser makes of it
(PyMIC/pymic/util/parse_config.py:86-111).  Build container only:  python -m oracle.gen_golden_cfg"""
import contextlib
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEXT = """[dataset]
tensor_type = float
task_type = seg
root_dir = ./data
1_train_csv = config/train_t1.csv
2_train_csv = config/train_t2_wi+wp.csv
test_csv = config/test_hrT2.csv
train_batch_size = 4
train_transform = [Pad, RandomCrop, RandomFlip, NormalizeWithMeanStd, LabelToProbability]
RandomCrop_output_size = [32, 128, 128]
RandomFlip_flip_depth = False
NormalizeWithMeanStd_channels = [0]

[network]
net_type = UNet2D5_dsbn
num_domains = 2
class_num = 2
in_chns = 1
feature_chns = [16, 32, 64, 128, 256]
conv_dims = [3, 3, 3, 3, 3]
dropout = [0.0, 0.0, 0.3, 0.4, 0.5]
bilinear = False
deep_supervise = False
aes = False

[training]
train_fpl_uda = True
dual = True
dis = False
val_t1 = False
val_t2 = True
gpus = [0]
loss_type = [DiceLoss, CrossEntropyLoss]
loss_weight = [0.5, 0.5]
optimizer = Adam
learning_rate = 1e-4
momentum = 0.9
weight_decay = 1e-5
lr_scheduler = MultiStepLR
lr_gamma = 0.5
lr_milestones = [10000, 20000, 30000]
ckpt_save_dir = model/vs_S
ckpt_save_prefix = vs
iter_start = 0
iter_max = 40000
iter_valid = 500
iter_save = 40000
early_stop_patience = None

[testing]
fpl = True
fpl_uncertainty_sorted = ./weight/sorted.npy
gpus = [0]
domian_label = 1
ae = False
ckpt_mode = 1
output_dir = result
evaluation_mode = True
test_time_dropout = True
tta_mode = 1
sliding_window_enable = True
sliding_window_size = [32, 128, 128]
sliding_window_stride = [32, 128, 128]
"""

if __name__ == "__main__":
    sys.path.insert(0, "/root/reference/PyMIC")
    from pymic.util.parse_config import parse_config, synchronize_config
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(TEXT)
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = synchronize_config(parse_config(f.name))
    os.unlink(f.name)
    with open(os.path.join(ROOT, "tests", "golden", "cfg_parse.json"), "w") as g:
        json.dump({"text": TEXT, "parsed": cfg}, g, indent=1)
    print("wrote cfg_parse.json with", sum(len(v) for v in cfg.values()), "keys")
