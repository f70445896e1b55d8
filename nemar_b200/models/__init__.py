"""--model registry (reference models/__init__.py:25-67): `--model foo` resolves to class FooModel in
nemar_b200/models/foo_model.py, a subclass of BaseModel."""
import importlib

from .base_model import BaseModel


def find_model_using_name(model_name):
    module = importlib.import_module("%s.%s_model" % (__name__, model_name))
    wanted = (model_name.replace("_", "") + "model").lower()
    for name, cls in vars(module).items():
        if name.lower() == wanted and isinstance(cls, type) and issubclass(cls, BaseModel):
            return cls
    raise NotImplementedError("%s_model.py must define a BaseModel subclass named like %s" % (model_name, wanted))


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(instance).__name__)
    return instance
