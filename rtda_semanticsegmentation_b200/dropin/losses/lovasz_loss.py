"""``losses/lovasz_loss.py`` of the reference, backed by the B200 kernels.

Same public names as the reference module (losses/lovasz_loss.py:14-200): ``isnan``,
``mean``, ``lovasz_grad``, ``LovaszSoftmax``; plus the Berman spellings.  Everything is
re-exported from ``rtda_semanticsegmentation_b200.lovasz_loss``.
"""
from rtda_semanticsegmentation_b200.lovasz_loss import (  # noqa: F401
    LovaszSoftmax, flatten_probas, isnan, lovasz_grad, lovasz_softmax, lovasz_softmax_flat, mean)
