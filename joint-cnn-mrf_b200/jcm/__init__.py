"""jcm - the joint-cnn-mrf hot path (part-detector CNN + MRF spatial model + heat-map loss) on hand-written sm_100a
CUDA kernels.  The functions re-exported here carry the names of the reference's main.py / evaluation.py."""
from ._lib import lib, JcmError, LIB_PATH  # noqa: F401
from . import ops  # noqa: F401
from .graph import (Context, PairwiseParams, JOINT_NAMES, model, spatial_model, conv_mrf, conv2d, batch_norm,  # noqa: F401
                    max_pool_layer, weight_variable, bias_variable, spatial_softmax, softmax_cross_entropy,
                    get_joints_coords, det_rate, get_pairwise_distr, init_part_detector, load_params, tower_forward,
                    n_filters, conv_specs, eval_error, conv_layer, weight_decay, average_gradients, grad_renorm,
                    params_updated)
from . import train  # noqa: F401,E402  (Trainer: the data-parallel training step, main.py:474-577)
from .feed import DeviceFeed  # noqa: F401,E402
from .checkpoint import save_checkpoint, load_checkpoint  # noqa: F401,E402
from . import augment  # noqa: F401,E402  (augmentation.py: augment_train / augment_test)
from . import dataset  # noqa: F401,E402  (data.py / prepare_pairwise_distribution.py: label, image and prior file formats)
from . import multiscale  # noqa: F401,E402  (main.py:326-425: multi-scale test-time inference)
