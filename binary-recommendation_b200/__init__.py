"""B200-native hot path of leotimus/binary-recommendation: NeuMF / BPR / two-tower training and
full-catalog top-K scoring, as hand-written sm_100a CUDA behind a C ABI (include/brk_b200.h).

Import as ``binrec_b200`` (this directory's name has a hyphen).  Modules mirror the reference:
  RModel, NeuMFModel, BPRModel, NCFModel   <- src/models/*.py
  SVD                            <- src/origin_models/svd/SVD.py
  twoTower, loadBinaryMovieLens  <- trainers/twoTower.py, trainers/loadBinaryMovieLens.py
  topKmetrics                    <- trainers/topKmetrics.py
  hotpath / _native              <- the Keras/TF ops underneath (gather, scatter, optimizers, ...)
"""
__version__ = "0.1.0"
