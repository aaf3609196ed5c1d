"""CPU: host-side logic of the product package that needs no GPU -- the C-ABI library loads and
exports every symbol include/mrgnas.h declares, module construction order / state_dict keys /
seeded init match the reference, and the product path refuses to run without CUDA."""
import ctypes
import os
import re
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")


def test_library_exports_every_declared_symbol():
    from mr_gnas_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "mrgnas.h")).read()
    declared = set(re.findall(r"\b(mrg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mrgnas.h but not exported"
    assert set(_lib.declared_symbols()) == declared
    lib.mrg_abi_version.restype = ctypes.c_int32
    assert lib.mrg_abi_version() == 1


def _args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D,
                                 conve_hid_drop=0.0, feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


def test_state_dict_keys_and_seeded_init_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "network_lp.pt"), weights_only=False)
    d = G["dims"]
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.utils import weights_init
    torch.manual_seed(0)
    np.random.seed(0)
    m = Network('cpu', eval(G["genotype"]), d["N"], d["R"], d["D"], d["D0"], 2 * d["R"] + 1, nn.BCELoss(), 0.0,
                _args(d["D"]))
    m.apply(weights_init)
    assert list(m.state_dict().keys()) == G["state_keys"]
    for k, v in m.state_dict().items():
        assert torch.equal(v, G["state0"][k]), k


def test_registries_match_reference_names():
    from mr_gnas_b200 import operations_lp as ops
    assert ops.PRE_OPS == ['pre_mult', 'pre_sub', 'pre_add']
    assert ops.FIRST_OPS == ['f_zero', 'f_identity', 'f_dense_comp', 'f_sparse_comp', 'f_comp']
    assert ops.MIDDLE_OPS == ['a_max', 'a_sum', 'a_mean']
    assert ops.LAST_OPS == ['f_zero', 'f_identity', 'f_dense_last', 'f_sparse_last']
    assert set(ops.MIXED_OPS) == {'pre_mult', 'pre_sub', 'pre_add', 'f_zero', 'f_identity', 'f_dense', 'f_dense_comp',
                                  'f_comp', 'f_sparse', 'f_sparse_comp', 'f_dense_last', 'f_sparse_last', 'a_max',
                                  'a_mean', 'a_sum'}
    assert set(ops.MIXED_OPS_sf) == {'sf_TransE', 'sf_DisMult', 'sf_ConvE'}


def test_product_path_has_no_cpu_fallback():
    from mr_gnas_b200 import functional as K
    with pytest.raises(RuntimeError):
        K.ComposeRows.apply(torch.randn(4, 8), torch.randn(4, 8), 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mr_gnas_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} references oracle/"


def test_supernet_construction_matches_reference(golden_dir):
    """LP and NC supernets: identical state_dict keys, seeded init (parameters and alpha tables) and decoded
    genotype strings (model_search_lp.py:215-311) -- all host-side logic."""
    from mr_gnas_b200.model_search import Network as SearchNC
    from mr_gnas_b200.model_search_lp import Network as SearchLP
    from mr_gnas_b200.utils import weights_init
    G = torch.load(os.path.join(golden_dir, "search_lp.pt"), weights_only=False)
    torch.manual_seed(5)
    m = SearchLP('cpu', G['num_ent'], G['num_rels'], 2, 1, 2, 2, G['D'], G['D0'], 2 * G['num_rels'] + 1, 40, 0.0, 0.0)
    m.apply(weights_init)
    assert list(m.state_dict().keys()) == G['state_keys']
    assert all(torch.equal(v, G['state0'][k]) for k, v in m.state_dict().items())
    assert all(torch.equal(a.detach(), b) for a, b in zip(m.arch_parameters(), G['alphas0']))
    assert str(m.show_genotypes()) == G['genotypes']
    G = torch.load(os.path.join(golden_dir, "network_nc.pt"), weights_only=False)
    r = G['search']
    torch.manual_seed(9)
    m = SearchNC('cpu', G['N'], G['C'], G['ET'], 2, 1, 2, G['D'], G['D0'], G['NB'], 0.0)
    m.apply(weights_init)
    assert list(m.state_dict().keys()) == r['state_keys']
    assert all(torch.equal(v, r['state0'][k]) for k, v in m.state_dict().items())
    assert str(m.show_genotypes()) == r['genotypes']


def test_nc_derived_network_keys_and_default_genotype(golden_dir):
    from mr_gnas_b200.genotypes import Genotype
    from mr_gnas_b200.model import Network
    G = torch.load(os.path.join(golden_dir, "network_nc.pt"), weights_only=False)
    geno = eval(G['genotype'])          # the NC default string has no score_func: must still parse
    assert geno[0].score_func is None
    m = Network('cpu', geno, G['N'], G['C'], G['ET'], 2, 1, 2, G['D'], G['D0'], G['NB'], nn.CrossEntropyLoss(),
                types.SimpleNamespace(feature_dim=G['D'], op_norm=True))
    assert list(m.state_dict().keys()) == G['derived_norm1']['state_keys']


def test_process_matches_reference_items(golden_dir):
    from mr_gnas_b200.process_data import make_batch, process
    G = torch.load(os.path.join(golden_dir, "network_lp.pt"), weights_only=False)
    gd = G["graph"]
    trip = gd["triples"].numpy()
    items = process({'train': trip, 'valid': trip[:5], 'test': trip[:5]}, gd["num_rels"])['train'][: G["dims"]["B"]]
    for it, (tr, lab) in zip(items, G["train_items"]):
        assert list(it["triple"]) == tr and sorted(it["label"]) == sorted(lab)
    t, y = make_batch(items, gd["num_ent"], lbl_smooth=0.1)
    assert torch.equal(y, G["labels"]) and torch.equal(t[:, 0], G["subj"]) and torch.equal(t[:, 1], G["rel"])


def test_sparse_batch_and_label_values_match_dense_host_labels():
    """Host side of the label pipeline: the CSR object lists + the two fp32 label values reproduce TrainDataset's
    dense smoothed rows (utils/data_set.py:17-33) bit for bit (the device kernel only places `pos` / `neg`)."""
    import numpy as np
    import torch
    from mr_gnas_b200.process_data import label_values, make_batch, make_batch_sparse, process
    from mr_gnas_b200.synth import synth_kg
    N, R = 977, 5
    trip = synth_kg(N, R, 3000, seed=3)
    items = process({'train': trip.tolist(), 'valid': [], 'test': []}, R)['train'][:64]
    for ls in (0.1, 0.0, 0.25):
        t, y = make_batch(items, N, lbl_smooth=ls)
        t2, ptr, idx = make_batch_sparse(items)
        assert torch.equal(t, t2) and ptr.dtype == torch.int32 and int(ptr[-1]) == idx.numel()
        neg, pos = label_values(N, ls)
        dense = torch.full((len(items), N), neg, dtype=torch.float32)
        for b in range(len(items)):
            dense[b, idx[ptr[b]:ptr[b + 1]].long()] = pos
        assert torch.equal(dense, y)
        # and the reference's own per-item formula
        y0 = np.zeros([N], dtype=np.float32)
        y0[np.int32(items[0]['label'])] = 1
        ref = torch.tensor(y0, dtype=torch.float32)
        if ls != 0.0:
            ref = (1.0 - ls) * ref + (1.0 / N)
        assert torch.equal(y[0], ref)


def test_batch_builders_match_real_reference_dataset(golden_dir):
    """process() + make_batch / make_batch_sparse + label_values against rows produced by the REAL
    utils/process_data.process and utils/data_set.TrainDataset (tests/golden/labels.pt)."""
    import torch
    from mr_gnas_b200.process_data import label_values, make_batch, make_batch_sparse, process
    G = torch.load(os.path.join(golden_dir, "labels.pt"), weights_only=False)
    N, R, trip = G["N"], G["R"], G["triples"].numpy()
    items = process({'train': trip, 'valid': trip[:0], 'test': trip[:0]}, R)['train'][:len(G["items"])]
    assert [it['triple'] for it in items] == [tuple(it['triple']) for it in G["items"]]
    assert [sorted(it['label']) for it in items] == [sorted(it['label']) for it in G["items"]]
    for ls, (t_ref, y_ref) in G["rows"].items():
        t, y = make_batch(items, N, lbl_smooth=ls)
        assert torch.equal(t, t_ref) and torch.equal(y, y_ref)
        t2, ptr, idx = make_batch_sparse(items)
        neg, pos = label_values(N, ls)
        dense = torch.full((len(items), N), neg, dtype=torch.float32)
        for b in range(len(items)):
            dense[b, idx[ptr[b]:ptr[b + 1]].long()] = pos
        assert torch.equal(dense, y_ref)


def test_shim_lets_the_reference_scripts_import_this_package():
    """mr_gnas_b200.shim: with the aliases installed the reference's own train/mr_lp_train.py and
    search/mr_lp_search.py import UNCHANGED (their `models.*`, `dgl`, `utils.process_data`, `utils.utils_rgcn`,
    `configs.genotypes` resolve to this package, `utils.utils` / `utils.data_set` / `models.architect_lp` to the
    reference's files) and their own helper code builds the network and the 1-N items.  Needs the reference
    checkout (build container only)."""
    import importlib.util
    import subprocess
    import sys
    ref = os.environ.get("MRG_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    code = r'''
import sys, importlib.util, types
import numpy as np, torch
from mr_gnas_b200 import shim
ref = shim.install(sys.argv[1])
for rel in ("train/mr_lp_train.py", "search/mr_lp_search.py"):
    spec = importlib.util.spec_from_file_location("ref_script", ref + "/" + rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                     # module scope only: imports + function definitions
    import mr_gnas_b200
    assert mod.Network.__module__.startswith("mr_gnas_b200."), mod.Network.__module__
    assert mod.process.__module__ == "mr_gnas_b200.process_data"
import models.operations_lp, models.architect_lp, utils.utils, configs.genotypes, dgl
assert models.operations_lp.__name__ == "mr_gnas_b200.operations_lp"
assert models.architect_lp.__file__.startswith(ref)          # the reference's own file
assert utils.utils.__file__.startswith(ref)
from mr_gnas_b200.graph import MRGraph
assert dgl.DGLGraph is MRGraph
# the script's own model construction path (mr_lp_train.py:101-129) on a small synthetic KG, on the host
spec = importlib.util.spec_from_file_location("ref_train", ref + "/train/mr_lp_train.py")
tr = importlib.util.module_from_spec(spec); spec.loader.exec_module(tr)
genotype = eval("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('a_max', 3, 2)], concat_node=[3], score_func='sf_DisMult')]",
                {"Genotype": configs.genotypes.Genotype})
args = types.SimpleNamespace(feature_dim=16, drop_aggr=0.1, drop_op=0.0, gamma=40, embed_dim=16, conve_hid_drop=0.3,
                             feat_drop=0.2, num_filt=4, ker_sz=3, k_w=4, k_h=4)
m = tr.Network('cpu', genotype, 50, 3, 16, 16, 7, torch.nn.BCELoss(), 0.3, args)
m.apply(tr.weights_init)
assert tr.count_parameters_in_MB(m) > 0
d = dgl.contrib.data.load_data("FB15k-237")
assert d.num_nodes == 14541 and d.train.shape == (272115, 3)
items = tr.process({'train': d.train[:500], 'valid': d.valid[:10], 'test': d.test[:10]}, d.num_rels)
ds = tr.TrainDataset(items['train'], d.num_nodes, types.SimpleNamespace(lbl_smooth=0.1))
t, y = ds[0]
assert y.shape == (14541,)
print("shim ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code, ref], capture_output=True, text=True, cwd=root,
                         env={**os.environ, "PYTHONPATH": root})
    assert out.returncode == 0 and "shim ok" in out.stdout, out.stderr[-2000:]
