"""Host-side mirror of the reference's ``torch_utils`` namespace, reduced to the operator hot path.

Put ``pasta-gan_b200/`` ahead of the reference on ``sys.path`` (or copy ``torch_utils/ops/*.py`` over the
reference's) and ``from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma``
resolves to the sm_100a implementation with the reference's call surface."""
