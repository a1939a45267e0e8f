import types, torch, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from realhf_b200.api.config import ModelName
from realhf_b200.api.model import FinetuneSpec, Model
from realhf_b200.engine.engine import TrainBackend
from realhf_b200.interfaces import basic
from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel
from test_parallel_cpu import _batch
cfg = hf_io.family("llama").make_test_config(); cfg.n_layers = 2
m = ReaLModel(cfg, dtype=torch.bfloat16, device=torch.device("cuda")).instantiate(seed=7)
tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant", grad_dtype="fp32", gradient_clipping=1.0)).initialize(Model(ModelName("m", 0), m, tok, "cuda"), FinetuneSpec(1, 10, 10))
itf = basic.SFTInterface()
for n_mbs in (1, 4):
    print(n_mbs, [round(itf.train_step(model, _batch(8).to_device("cuda"), n_mbs=n_mbs)["loss"], 4) for _ in range(3)])
