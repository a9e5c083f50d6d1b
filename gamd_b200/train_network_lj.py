"""Drop-in for the reference's train_network_lj module (inference surface only): same module-level
constants and ParticleNetLightning entry point; training is out of scope."""
from .force_field import LJForceFacade, create_water_bond  # noqa: F401

CUTOFF_RADIUS = 7.5
BOX_SIZE = 27.27
NUM_OF_ATOMS = 258


class ParticleNetLightning(LJForceFacade):
    def __init__(self, args, num_device=1, epoch_num=100, batch_size=1, learning_rate=3e-4, log_freq=1000,
                 model_weights_ckpt=None, scaler_ckpt=None, **kw):
        super().__init__(args, BOX_SIZE, CUTOFF_RADIUS, NUM_OF_ATOMS, model_weights_ckpt, scaler_ckpt, **kw)
        self.cutoff = CUTOFF_RADIUS
