"""
Drop-in aliases: after `dlwp_b200.compat.install()` the statements the reference's scripts start with keep working
unchanged (examples/train.py:15-21, examples/train_functional.py:19-28):

    from DLWP.model import DLWPNeuralNet, DLWPFunctional
    from DLWP.custom import PeriodicPadding2D, RowConnected2D, slice_layer, EarlyStoppingMin, ...
    from DLWP.util import save_model, load_model
    from keras.layers import Input, Conv2D, ZeroPadding2D, MaxPooling2D, UpSampling2D, concatenate, ...
    from keras.models import Model
"""

import sys
import types


def install(force=False):
    import dlwp_b200
    from . import custom, keras, model, util
    from .keras.layers import convolutional, local
    from .model import generators, models
    if 'keras' in sys.modules and not force and not getattr(sys.modules['keras'], '__version__', '').endswith(
            'dlwp_b200'):
        raise RuntimeError('a real `keras` is already imported; pass force=True to shadow it with dlwp_b200.keras')
    pkg = types.ModuleType('DLWP')
    pkg.__path__ = []
    pkg.__version__ = '0.7.0+dlwp_b200'
    pkg.model, pkg.custom, pkg.util = model, custom, util
    mods = {
        'DLWP': pkg, 'DLWP.model': model, 'DLWP.model.models': models, 'DLWP.model.generators': generators,
        'DLWP.custom': custom, 'DLWP.util': util,
        'keras': keras, 'keras.layers': keras.layers, 'keras.layers.convolutional': convolutional,
        'keras.layers.local': local, 'keras.models': keras.models, 'keras.callbacks': keras.callbacks,
        'keras.regularizers': keras.regularizers, 'keras.losses': keras.losses, 'keras.backend': keras.backend,
        'keras.utils': keras.utils, 'keras.optimizers': keras.optimizers,
    }
    # every other submodule under its reference name too (`DLWP.model.extensions`, `DLWP.model.models_torch`, `keras.engine`,
    # ...): without the alias the import system would find the file through the aliased package's __path__ and execute it
    # a SECOND time under the new name -- duplicate classes that fail isinstance checks (and relative imports that resolve
    # against `DLWP` instead of `dlwp_b200`)
    import importlib
    import pkgutil
    for package, alias in ((model, 'DLWP.model'), (keras, 'keras')):
        for info in pkgutil.walk_packages(package.__path__, package.__name__ + '.'):
            sub = importlib.import_module(info.name)
            mods.setdefault(alias + info.name[len(package.__name__):], sub)
    sys.modules.update(mods)
    return dlwp_b200
