"""Only the helpers the hypercomplex path and the training scripts import from phc.quaternion.
The quaternion model family itself is out of scope."""
