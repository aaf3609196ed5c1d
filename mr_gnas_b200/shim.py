"""Run the reference's OWN scripts unchanged on top of this package.

    python -m mr_gnas_b200.shim /path/to/MR-GNAS/train/mr_lp_train.py --device cuda:0 --dataset FB15k-237 ...

`install(reference_root)` puts module aliases into ``sys.modules`` so that the scripts' imports resolve to the
B200-native implementations for everything on the message-passing path and to the reference's own files for the
rest (argument parsing, logging, training loop, DARTS architect, metrics, datasets):

  models.operations_lp / operations / model_lp / model / model_search_lp / model_search / cell_lp / cell / compgcn
                                  -> mr_gnas_b200.<same name>            (models.architect*: the reference's files)
  configs.genotypes               -> mr_gnas_b200.genotypes             (Genotype with the score_func=None default)
  utils.process_data, utils.utils_rgcn -> mr_gnas_b200.process_data / utils_rgcn (utils.utils, utils.data_set: reference)
  utils.gpu_memory_log            -> a no-op (the file is missing from the reference repository)
  dgl                             -> a module whose DGLGraph / graph() build mr_gnas_b200.graph.MRGraph (no DGL needed)
  dgl.contrib.data.load_data, dataloader.get_dataset, dgl.data.rdf.*  -> SYNTHETIC datasets of the published shapes
                                     (there is no network on the build / GPU machines; pass real loaders to install()
                                      to train on the real data)
  tensorboardX.SummaryWriter      -> torch.utils.tensorboard's writer when importable, else a no-op

Nothing here is on the measured path; it is integration glue."""
import importlib
import os
import runpy
import sys
import types

_OURS = ["operations_lp", "operations", "model_lp", "model", "model_search_lp", "model_search", "cell_lp", "cell",
         "compgcn"]


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _SyntheticKG:
    """What dgl.contrib.data.load_data('FB15k-237') returns, with synthetic triples of that shape."""

    def __init__(self, num_nodes, num_rels, n_train, n_valid, n_test, seed=0):
        from .synth import synth_kg
        t = synth_kg(num_nodes, num_rels, n_train + n_valid + n_test, seed=seed)
        self.num_nodes, self.num_rels = num_nodes, num_rels
        self.train, self.valid, self.test = t[:n_train], t[n_train:n_train + n_valid], t[n_train + n_valid:]
        # dataloader.get_dataset(...) flavour: columns-first lists
        self.n_entities, self.n_relations = num_nodes, num_rels


def _load_data(name, *a, **k):
    if str(name).lower().startswith("fb15k"):
        return _SyntheticKG(14541, 237, 272115, 17535, 20466)
    return _SyntheticKG(40943, 11, 86835, 3034, 3134)


def _get_dataset(data_path, data_name, format_str, files=None):
    d = _load_data(data_name)
    d.train, d.valid, d.test = (x.transpose().tolist() for x in (d.train, d.valid, d.test))   # scripts transpose back
    return d


class _NoWriter:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def install(reference_root=None, load_data=None, get_dataset=None):
    """Install the aliases.  `reference_root`: checkout of Amanda-Zheng/MR-GNAS (for the files that stay the
    reference's own); `load_data` / `get_dataset`: real dataset loaders to use instead of the synthetic ones."""
    from . import genotypes, process_data, utils_rgcn
    from .graph import MRGraph
    ref = os.path.abspath(reference_root) if reference_root else None
    models = _module("models", __path__=[os.path.join(ref, "models")] if ref else [])
    for name in _OURS:
        mod = importlib.import_module(f"mr_gnas_b200.{name}")
        sys.modules[f"models.{name}"] = mod
        setattr(models, name, mod)
    configs = _module("configs", __path__=[os.path.join(ref, "configs")] if ref else [])
    sys.modules["configs.genotypes"] = genotypes
    configs.genotypes = genotypes
    utils = _module("utils", __path__=[os.path.join(ref, "utils")] if ref else [])
    sys.modules["utils.process_data"] = process_data
    sys.modules["utils.utils_rgcn"] = utils_rgcn
    utils.process_data, utils.utils_rgcn = process_data, utils_rgcn
    utils.gpu_memory_log = _module("utils.gpu_memory_log", gpu_memory_log=lambda *a, **k: None)
    # ---- dgl
    dgl = _module("dgl", __path__=[], DGLGraph=MRGraph, graph=lambda data=None, **k: MRGraph(), EID="_ID", NID="_ID",
                  ETYPE="_TYPE")
    desc = lambda kind: (lambda *a, **k: (kind,) + a)
    dgl.function = _module("dgl.function", copy_edge=desc("copy"), copy_e=desc("copy"), max=desc("max"),
                           sum=desc("sum"), mean=desc("mean"), u_sub_e=desc("u_sub_e"), u_mul_e=desc("u_mul_e"))
    dgl.contrib = _module("dgl.contrib", __path__=[])
    dgl.contrib.data = _module("dgl.contrib.data", load_data=load_data or _load_data)
    dgl.data = _module("dgl.data", __path__=[])
    dgl.data.rdf = _module("dgl.data.rdf", AIFBDataset=None, MUTAGDataset=None, BGSDataset=None, AMDataset=None)
    _module("dataloader", get_dataset=get_dataset or _get_dataset)
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception:
        SummaryWriter = _NoWriter
    _module("tensorboardX", SummaryWriter=SummaryWriter)
    if ref and ref not in sys.path:
        sys.path.insert(0, ref)
    return ref


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    root = os.path.dirname(os.path.dirname(script))
    install(root)
    os.chdir(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
