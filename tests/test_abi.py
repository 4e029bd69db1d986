"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the
header declares; the host modules expose the reference's names and state_dict keys; the product
package never touches the oracle; CUDA-only ops refuse CPU tensors."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "mirage_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mirage_b200 import _lib, build
    build.build()
    handle = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/mirage_b200.h but not exported"
    assert set(declared) == set(_lib.exported_symbols())
    assert handle.mb_version() == 1


def test_product_never_imports_oracle():
    for f in (ROOT / "mirage_b200").glob("*.py"):
        src = f.read_text()
        assert "oracle" not in src.replace("oracle/", ""), f"{f.name} references the oracle"


def test_ops_refuse_cpu_tensors():
    from mirage_b200 import _lib, ops
    x = torch.randn(8, 128)
    with pytest.raises(_lib.MirageB200Error):
        ops.layernorm(x, torch.ones(128), torch.zeros(128))


def test_hf_wrapper_api_and_state_dict_keys():
    from mirage_b200.mirage_hf import MIRAGEWrapper
    m = MIRAGEWrapper(input_size=512, patch_size=32, modalities="bscan-slo", size="base")
    sd = m.model.state_dict()
    assert sd["global_tokens"].shape == (1, 1, 768)
    assert sd["input_adapters.bscan.proj.weight"].shape == (768, 1, 32, 32)
    assert sd["input_adapters.slo.pos_emb"].shape == (1, 768, 16, 16)
    assert not m.model.input_adapters["bscan"].pos_emb.requires_grad
    assert sd["encoder.11.attn.qkv.weight"].shape == (2304, 768)
    assert sd["encoder.0.mlp.fc1.weight"].shape == (3072, 768)
    assert len(sd) == 151 and sum(v.numel() for v in sd.values()) == 87022848
    assert m.model.output_adapters is None and len(m.model.encoder) == 12
    with pytest.raises(ValueError):
        MIRAGEWrapper(size="huge")
    # load_state_dict forwards to .model (keys without the "model." prefix)
    m.load_state_dict(sd)


def test_pretrain_model_parameter_inventory():
    from pretrain_case import build_pretrain_model
    base, _ = build_pretrain_model("base")
    assert sum(p.numel() for p in base.parameters() if p.requires_grad) == 98220160
    frozen = [k for k, p in base.named_parameters() if not p.requires_grad]
    assert len(frozen) == 6 and all(k.endswith("pos_emb") for k in frozen)
    assert base.no_weight_decay() >= {"global_tokens", "input_adapters.bscan.pos_emb",
                                      "input_adapters.bscanlayermap.class_emb",
                                      "output_adapters.slo.mask_token"}
    assert base.get_num_layers() == 12


def test_factories_and_adapter_errors():
    from mirage_b200.input_adapters import PatchedInputAdapter
    from mirage_b200.model import model_factory
    assert {"miragepre_base", "miragepre_large", "miragelight_base", "miragelight_large"} <= set(model_factory)
    ad = PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32), image_size=512)
    with pytest.raises(AssertionError, match="init"):
        ad(torch.zeros(1, 1, 512, 512))


def test_masks_host_logic_bit_exact_with_reference_golden():
    """generate_random_masks on the CPU generator reproduces the reference's recorded outputs."""
    from helpers import GOLDEN
    from pretrain_case import build_pretrain_model
    model, _ = build_pretrain_model("tiny")
    g = torch.load(GOLDEN / "masks.pt")
    for case in g["cases"]:
        torch.manual_seed(case["seed"])
        toks = {d: torch.empty(case["B"], 256, 0) for d in g["domains"]}
        tm, keep, restore = model.generate_random_masks(toks, case["n_vis"], alphas=case["alphas"])
        assert torch.equal(keep, case["ids_keep"].long())
        assert torch.equal(restore, case["ids_restore"].long())
        for d in g["domains"]:
            assert torch.equal(tm[d], case["task_masks"][d].long())
