"""Base class of the recommender models -- the drop-in mirror of the reference's
src/models/RModel.py (same attribute names, property setters, train() orchestration and return
value), with the Keras model underneath replaced by device-resident tables driven through the
C ABI.

Deliberate differences from the reference (SURVEY.md section 0.5 lists its broken entry points):
  * `train` works without a distributedConfig (the reference raises NameError at RModel.py:139
    because `strategy` is unbound);
  * data may be given as a CSV path with the reference's columns (NeuMFModel.py:21-27), as a ".brkc"
    binary columnar cache of it (interactions.py), or directly as a (users, items) pair of int arrays /
    a dict with CUSTOMER_ID and PRODUCT_ID;
  * splits, shuffles and samplers are seeded (the reference's are not reproducible);
  * plotting / model_to_dot / SMB access are out of scope (not compute).
"""
import os

import numpy as np

from . import synth


class LocalDataStore:
    """Local-file half of the reference's DataStore.openFile (src/datasource/DataStore.py:12-16);
    the SMB half needs credentials and a network and is out of scope."""

    def openFile(self, path, mode="r"):
        return open(path, mode)


class RModel:
    CUSTOMER_ID = 'CUSTOMER_ID'
    PRODUCT_ID = 'PRODUCT_ID'
    METRICS = ['mse', 'mae', 'binary_accuracy']

    def __init__(self, moduleName, workDir=None):
        self.modelName = moduleName
        base = workDir or os.getcwd()
        self.checkpointPath = os.path.join(base, 'checkpoints/{}/cp'.format(self.modelName))
        self.modelProducts = os.path.join(base, 'checkpoints/{}/modelData/products'.format(self.modelName))
        self.modelUsers = os.path.join(base, 'checkpoints/{}/modelData/users'.format(self.modelName))

        self.numFactor: int = 32          # RModel.py:35
        self._epochs: int = 10            # RModel.py:36
        self._batchSize: int = 1024       # RModel.py:37
        self._validationSteps: int = 20   # RModel.py:38
        self._testSize: float = 0.2       # RModel.py:39
        self._model = None
        self._dataStore = LocalDataStore()
        self.seed = 42                    # weights (bpr.py:21 uses SEED = 42)
        self.samplerSeed = 7
        self.splitSeed = synth.DATA_SEED + 1
        self.history = None

    # -- properties, as RModel.py:45-91 ------------------------------------------------------
    @property
    def testSize(self) -> float:
        return self._testSize

    @testSize.setter
    def testSize(self, value: float):
        self._testSize = value

    @property
    def validationSteps(self) -> int:
        return self._validationSteps

    @validationSteps.setter
    def validationSteps(self, value):
        self._validationSteps = value

    @property
    def epochs(self) -> int:
        return self._epochs

    @epochs.setter
    def epochs(self, value: int):
        self._epochs = value

    @property
    def batchSize(self) -> int:
        return self._batchSize

    @batchSize.setter
    def batchSize(self, value: int):
        self._batchSize = value

    @property
    def dataStore(self):
        return self._dataStore

    @dataStore.setter
    def dataStore(self, value):
        pass                               # read-only, as in the reference (RModel.py:82-84)

    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, value):
        self._model = value

    # -- to be specialised ---------------------------------------------------------------------
    def compileModel(self, distributedConfig, numUser: int, numItem: int, numFactor: int):
        print('placeholder')
        return None

    def bootstrapDataset(self, df, negRatio=3., batchSize=128, shuffle=True):
        return None

    def prepareToTrain(self, distributedConfig, path, rowLimit):
        return None

    def predictForUser(self, customerId, numberOfItem=5):
        return None

    def getPredictableUsers(self) -> list:
        return []

    def getPredictDataFrame(self, customerId):
        return None

    def plot(self, history, metrics):
        """RModel.py:100-113 draws loss / val_loss / the given metrics per epoch into trainResultPlotPath.  Plotting is
        outside the hot path: the same series are written as JSON beside the checkpoint (and drawn when matplotlib is
        importable).  history: object with .history (Keras) or a dict of lists; metrics: iterable of (key, label)."""
        import json
        series = getattr(history, "history", history) or {}
        keep = {k: [float(x) for x in v] for k, v in series.items()
                if k in ("loss", "val_loss") or k in {m for m, _ in (metrics.items() if isinstance(metrics, dict) else metrics)}}
        path = os.path.join(os.path.dirname(self.checkpointPath), "trainResult.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(keep, f)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return path
        fig, ax = plt.subplots(nrows=1, ncols=1)
        for k, v in keep.items():
            ax.plot(v, label=k)
        plt.ylim([0, 1]); plt.xlabel('Epoch'); plt.ylabel('Error'); plt.legend()
        fig.savefig(os.path.splitext(path)[0] + ".png")
        plt.close(fig)
        return path

    def train(self, path, rowLimit, metricDict: dict = {}, distributedConfig=None):
        """RModel.py:115-150, the generic flow a model class inherits: prepareToTrain -> fit (validation on the test
        split; steps_per_epoch = len / epochs / workers when distributed) -> save -> plot -> evaluate on a 20 % split
        of the training pairs -> {'result': 'completed', 'metrics': [...]}.  Under torchrun the data-parallel paths are
        picked up by the model itself (no strategy scope is needed)."""
        from . import synth
        trainDataset, testDataset, trainSplit = self.prepareToTrain(distributedConfig, path, rowLimit)
        if distributedConfig is None:
            history = self.model.fit(trainDataset, validation_data=testDataset, epochs=self.epochs)
        else:
            steps = int(len(trainDataset) / self.epochs / self.getNumberOfWorkers(distributedConfig))
            history = self.model.fit(trainDataset, validation_data=testDataset, epochs=self.epochs,
                                     steps_per_epoch=max(steps, 1))
        self.saveCheckPoint()
        self.plot(history, metricDict)
        print("Evaluating trained model...")
        _, val = synth.train_test_split(trainSplit[0], trainSplit[1], 0.2, seed=self.splitSeed + 1)
        valDataset = self.bootstrapDataset(val, shuffle=False)
        evaluatedMetric = list(self.model.evaluate(valDataset, steps=self.validationSteps))
        return {'result': 'completed', 'metrics': evaluatedMetric}

    def readyToTrain(self):
        return True

    def getNumberOfWorkers(self, distributedConfig) -> int:
        """Reference: len(TF_CONFIG['cluster']['worker']) (RModel.py:201-202).  Here a
        distributedConfig may be that dict or None; with torch.distributed initialised the world
        size wins."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_world_size()
        except Exception:
            pass
        if distributedConfig is None:
            return 1
        return len(distributedConfig['cluster']['worker'])

    # -- data --------------------------------------------------------------------------------
    def readData(self, path, rowLimit):
        """Returns (numItem, numUser, (users, items)) -- NeuMFModel.py:21-27: the table sizes are
        max id + 1."""
        users, items = self._loadPairs(path, rowLimit)
        numUser = int(users.max()) + 1
        numItem = int(items.max()) + 1
        return numItem, numUser, (users, items)

    def _loadPairs(self, path, rowLimit):
        if isinstance(path, dict):
            users, items = path[self.CUSTOMER_ID], path[self.PRODUCT_ID]
        elif isinstance(path, (tuple, list)) and len(path) == 2:
            users, items = path
        elif isinstance(path, str) and path.endswith(".brkc"):
            # binary columnar cache written by interactions.csv_to_cache: memory-mapped columns
            from .interactions import InteractionCache
            cache = InteractionCache(path)
            users, items = cache.columns["user"], cache.columns["item"]
        else:
            import pandas as pd
            with self.dataStore.openFile(path=path, mode='r') as f:
                df = pd.read_csv(f, nrows=rowLimit)
            users, items = df[self.CUSTOMER_ID].values, df[self.PRODUCT_ID].values
        users = np.ascontiguousarray(np.asarray(users)[:rowLimit], dtype=np.int32)
        items = np.ascontiguousarray(np.asarray(items)[:rowLimit], dtype=np.int32)
        if len(users) != len(items) or len(users) == 0:
            raise ValueError("need equally long, non-empty CUSTOMER_ID / PRODUCT_ID columns")
        if users.min() < 0 or items.min() < 0:
            raise ValueError("ids must be non-negative")
        return users, items

    # -- checkpoint (RModel.py:139,172-196) ----------------------------------------------------
    def isMaster(self, taskType, taskId) -> bool:
        return taskType is None or taskType == 'chief' or (taskType == 'worker' and taskId == 0)

    def getModelSaveLocation(self, strategy) -> str:
        """RModel.py:175-179: the chief saves into checkpointPath, every other worker into a temp dir that is removed
        afterwards (Keras needs all workers to call save).  `strategy` is None (single process) or anything exposing
        cluster_resolver.task_type / task_id like tf.distribute's strategies.  Kept for callers of the reference API;
        saveCheckPoint itself needs no temp dirs (every rank writes its own shard files, checkpoint.py)."""
        if strategy is None or self.isMaster(strategy.cluster_resolver.task_type, strategy.cluster_resolver.task_id):
            return self.checkpointPath
        return self.getSlaveTempDir(strategy.cluster_resolver.task_id)

    def getSlaveTempDir(self, taskId):
        """RModel.py:187-191."""
        tempDir = os.path.join(self.checkpointPath, 'workertemp_' + str(taskId))
        os.makedirs(tempDir, exist_ok=True)
        return tempDir

    def clearSlaveTempDir(self, strategy):
        """RModel.py:193-196 (the reference removes dirname(tempDir), i.e. the whole checkpoint directory of that
        worker's file system view; here only the worker's own temp dir is removed)."""
        import shutil
        if strategy is not None and self.isMaster(strategy.cluster_resolver.task_type,
                                                  strategy.cluster_resolver.task_id) is False:
            shutil.rmtree(self.getSlaveTempDir(strategy.cluster_resolver.task_id), ignore_errors=True)

    def getPredictDataSet(self, customerId):
        """RModel.py:168-170."""
        return self.bootstrapDataset(self.getPredictDataFrame(customerId), shuffle=False)

    def checkpointMeta(self) -> dict:
        """What restoreFromLatestCheckPoint needs to rebuild the model before loading tensors (the reference's
        SavedModel carries its graph; its test users / products sit in the pickles of RModel.py:28-29)."""
        return {"model": self.modelName}

    def buildFromMeta(self, meta: dict):
        raise NotImplementedError

    def saveCheckPoint(self):
        """model.save(checkpointPath) (RModel.py:139): a checkpoint directory (checkpoint.py) holding the weights,
        optimizer slots and step -- readable by a process with any number of GPUs."""
        from . import checkpoint as CK
        from . import distributed as D
        sd = self.model.state_dict()                 # collective under mirrored data parallelism (sharded Adam moments)
        if D.rank() == 0:                            # replicas are identical: the chief saves (RModel.py:175-196)
            CK.save_state_dict(self.checkpointPath, sd, meta=self.checkpointMeta())
        D.barrier()

    def restoreFromLatestCheckPoint(self):
        """tf.keras.models.load_model(checkpointPath) (RModel.py:172-173), as the REST endpoint calls it on a fresh
        object (src/restful/RecommendationEndpoint.py:49): the model is rebuilt from the sizes recorded in the
        checkpoint, then the tensors are loaded."""
        from . import checkpoint as CK
        meta = CK.read_manifest(self.checkpointPath).get("meta", {})
        if self.model is None:
            self.buildFromMeta(meta)
        self.model.load_state_dict(CK.load_state_dict(self.checkpointPath))
