"""reference ``phc.quaternion`` import paths.  The quaternion model family runs on the PHM kernels as n = 4 with a frozen
Hamilton rule (phc_gnn_b200/quaternion.py); the ``QTensor`` algebra of the reference is not needed on that path and is not
provided."""
