"""Import the reference's own pure-torch building blocks from /root/reference (read-only).

Test infrastructure only; usable ONLY in the build container (the GPU box has no
/root/reference).  The reference's packages pull in ``ride``, ``pytorch_lightning`` and
``continual`` at import time (datasets/datasets.py:4-7, models/base.py:6,10-11), none of which is
installed, so their package ``__init__`` files are bypassed with namespace modules and three
import-only stubs are registered (recipe: SURVEY.md Appendix A).  What comes out is the reference's
real ``GraphConvolution`` / ``AdaptiveGraphConvolution`` / ``TemporalConvolution`` /
``SpatioTemporalBlock`` / ``init_weights`` / ``graph.A``; the ``Co*`` factories are NOT usable (they need the real library).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("COSK_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "base.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference classes.  Idempotent."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "models.base" in sys.modules and getattr(sys.modules["models.base"], "_cosk_shim", False):
        return _namespace()
    import logging

    import torch.nn as nn

    class _Configs:
        def __init__(self):
            self.names = []

        def add(self, name=None, **kw):
            self.names.append(name)

    class _RideMixin:
        pass

    # class-statement-only stand-ins for what models/a_gcn/a_gcn.py:72-78 lists as bases of AGcn
    ride = _stub("ride", getLogger=logging.getLogger, Configs=_Configs, RideModule=type("RideModule", (), {}),
                 TopKAccuracyMetric=lambda *k: type("TopKAccuracyMetric", (), {}),
                 SgdOneCycleOptimizer=type("SgdOneCycleOptimizer", (), {}))
    ride.finetune = _stub("ride.finetune", Finetunable=type("Finetunable", (), {}))
    _stub("ride.core", Configs=_Configs, RideMixin=_RideMixin)
    _stub("ride.logging", getLogger=logging.getLogger)
    _stub("continual", Sequential=nn.Sequential)
    for pkg in ("datasets", "models"):
        ns = types.ModuleType(pkg)
        ns.__path__ = [os.path.join(REF_ROOT, pkg)]
        sys.modules[pkg] = ns
    for mod in ("datasets.graph", "datasets.ntu_rgbd", "datasets.kinetics", "models.utils", "models.base"):
        importlib.import_module(mod)
    # datasets/datasets.py needs ride + pytorch_lightning; a_gcn.py only names datasets.GraphDatasets as a base
    _stub("datasets.datasets", GraphDatasets=type("GraphDatasets", (), {}))
    sys.modules["datasets"].datasets = sys.modules["datasets.datasets"]
    ns = types.ModuleType("models.a_gcn")
    ns.__path__ = [os.path.join(REF_ROOT, "models", "a_gcn")]
    sys.modules["models.a_gcn"] = ns
    importlib.import_module("models.a_gcn.a_gcn")
    # models/s_tr/s_tr.py:16 pulls the training optimizer mixin; only the class statement of STr needs the name
    _stub("optimizers", SgdMultiStepLR=type("SgdMultiStepLR", (), {}))
    ns = types.ModuleType("models.s_tr")
    ns.__path__ = [os.path.join(REF_ROOT, "models", "s_tr")]
    sys.modules["models.s_tr"] = ns
    importlib.import_module("models.s_tr.s_tr")
    sys.modules["models.base"]._cosk_shim = True
    return _namespace()


def _namespace():
    base, utils = sys.modules["models.base"], sys.modules["models.utils"]
    return types.SimpleNamespace(
        GraphConvolution=base.GraphConvolution,
        TemporalConvolution=base.TemporalConvolution,
        SpatioTemporalBlock=base.SpatioTemporalBlock,
        AdaptiveGraphConvolution=sys.modules["models.a_gcn.a_gcn"].AdaptiveGraphConvolution,
        GcnUnitAttention=sys.modules["models.s_tr.s_tr"].GcnUnitAttention,
        init_weights=utils.init_weights,
        ntu_A=sys.modules["datasets.ntu_rgbd"].graph.A,
        kinetics_A=sys.modules["datasets.kinetics"].graph.A,
    )
