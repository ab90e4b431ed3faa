"""`get_model(name, **kwargs)`: the reference's single entry point (pytorchcv/model_provider.py:1364-1382).

Lower-cases the name, raises ValueError("Unsupported model: ...") for names outside the registry, forwards kwargs to
the constructor.  The registry holds every named constructor whose blocks are on the B200 eval path.
"""
from __future__ import annotations

from . import nets

__all__ = ["get_model", "supported_models"]

_models = {}


def _register(name, fn):
    fn.__name__ = name
    _models[name] = fn


for _name, _fixed in nets.RESNET_VARIANTS.items():
    _register(_name, nets._resnet_ctor(_name, **_fixed))
for _name, _fixed in nets.MOBILENETV2_VARIANTS.items():
    _register(_name, (lambda n, f: lambda **kw: nets.get_mobilenetv2(model_name=n, **f, **kw))(_name, _fixed))
for _name, _ws in nets.MOBILENET_VARIANTS.items():
    _register(_name, (lambda n, w: lambda **kw: nets.get_mobilenet(width_scale=w, model_name=n, **kw))(_name, _ws))
for _name, (_b, _c, _w) in nets.SERESNEXT_VARIANTS.items():
    _register(_name, (lambda n, b, c, w: lambda **kw: nets.get_seresnext(
        blocks=b, cardinality=c, bottleneck_width=w, model_name=n, **kw))(_name, _b, _c, _w))
for _name, (_b, _c, _w) in nets.RESNEXT_VARIANTS.items():
    _register(_name, (lambda n, b, c, w: lambda **kw: nets.get_resnext(
        blocks=b, cardinality=c, bottleneck_width=w, model_name=n, **kw))(_name, _b, _c, _w))
for _name, _b in nets.RESNETD_VARIANTS.items():
    _register(_name, (lambda n, b: lambda **kw: nets.get_resnetd(
        blocks=b, conv1_stride=False, model_name=n, **kw))(_name, _b))
for _name, (_v, _sz, _tf, _eps) in nets.EFFICIENTNET_VARIANTS.items():
    _register(_name, (lambda n, v, sz, tf, eps: lambda in_size=None, **kw: nets.get_efficientnet(
        version=v, in_size=in_size if in_size is not None else (sz, sz), tf_mode=tf, bn_eps=eps, model_name=n, **kw))(
        _name, _v, _sz, _tf, _eps))
for _name, (_ver, _ws) in nets.MOBILENETV3_VARIANTS.items():
    _register(_name, (lambda n, v, w: lambda **kw: nets.get_mobilenetv3(version=v, width_scale=w, model_name=n, **kw))(
        _name, _ver, _ws))
for _name, (_b, _k) in nets.DEEPLABV3_VARIANTS.items():
    _register(_name, nets._deeplab_ctor(_name, _b, _k))
for _name, (_b, _k) in nets.FCN8SD_VARIANTS.items():
    _register(_name, nets._fcn8sd_ctor(_name, _b, _k))
for _name, _ver in nets.PROXYLESSNAS_VARIANTS.items():
    _register(_name, (lambda n, v: lambda **kw: nets.get_proxylessnas(version=v, model_name=n, **kw))(_name, _ver))
_register("spnasnet", lambda **kw: nets.get_spnasnet(model_name="spnasnet", **kw))
for _name, _fixed in nets.FBNET_VARIANTS.items():
    _register(_name, (lambda n, f: lambda **kw: nets.get_fbnet(model_name=n, **f, **kw))(_name, _fixed))
for _name, _ver in nets.MNASNET_VARIANTS.items():
    _register(_name, (lambda n, v: lambda **kw: nets.get_mnasnet(version=v, width_scale=1.0, model_name=n, **kw))(
        _name, _ver))
for _name, (_b, _k) in nets.PSPNET_VARIANTS.items():
    _register(_name, nets._pspnet_ctor(_name, _b, _k))
for _name, _blocks in nets.SENET_VARIANTS.items():
    _register(_name, (lambda n, b: lambda **kw: nets.get_senet(blocks=b, model_name=n, **kw))(_name, _blocks))
for _name, _fixed in nets.SERESNET_VARIANTS.items():
    _register(_name, (lambda n, f: lambda **kw: nets.get_seresnet(model_name=n, **f, **kw))(_name, _fixed))


def supported_models() -> list[str]:
    return sorted(_models)


def get_model(name, **kwargs):
    """Get supported model: same contract as the reference (model_provider.py:1364-1382)."""
    name = name.lower()
    if name not in _models:
        raise ValueError("Unsupported model: {}".format(name))
    return _models[name](**kwargs)
