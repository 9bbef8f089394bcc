"""Caption-model factory -- drop-in for lib/caption_models/__init__.py:setup (att2in2 only).

The lang2seg training scripts only ever select `att2in2` (README.md:59,72; the pickled options of
every shipped caption_log_*); the other model families of self-critical.pytorch are out of scope
 and raise here.
"""
from .AttModel import Att2in2Core, Att2in2Model, Attention, AttModel  # noqa: F401


def setup(opt):
    if opt["caption_model"] != "att2in2":
        raise Exception("Caption model not supported on this path: {}".format(opt["caption_model"]))
    return Att2in2Model(opt)
