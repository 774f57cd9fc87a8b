"""NCFModel -- drop-in mirror of the reference's serve-only class (src/models/NCFModel.py): the NeuMF network of the
script trainers/NFC_plain.py (latent 10, Dense 100-50-10 sigmoid, BCE, head [pred_mf, pred_mlp]; :109-155) restored
from a checkpoint and queried per user over ALL product ids of its training file.

The reference's predictForUser (:42-51) runs model.predict over one (customer, product) frame, drops predictions
>= 1.0, sorts the rest descending and returns {product: '%.9f' % score} for the best numberOfItem.  Here the frame is
one fused-forward launch over the catalog (csrc/neumf2.cu, inference mode) and the selection a device top-k
(predictions >= 1.0 masked out first); several users go through predictForUsers in one pass.
"""
import numpy as np
import torch

from . import hotpath as H
from .NeuMFModel import NeuMFNet
from .RModel import RModel

SCRIPT_SPEC = dict(numFactor=10, hidden=(100, 50, 10), act="sigmoid", loss="bce", learning_rate=0.005,
                   head_order="mf_h3")                                  # NFC_plain.py:109-155


class NCFModel(RModel):
    def __init__(self, workDir=None, trainData=None):
        super().__init__('NCFModel', workDir)
        self._productIds: list = []
        self._customerIds: list = []
        # the reference hard-codes an SMB path (NCFModel.py:13) and loads it in the constructor; here the file is
        # optional at construction (a checkpoint carries both id lists) and any local CSV / .brkc cache is accepted
        self.trainData = trainData
        if trainData is not None:
            self.loadTrainData()

    @property
    def customerIds(self) -> list:
        return self._customerIds

    @customerIds.setter
    def customerIds(self, ids: list):
        self._customerIds = ids

    @property
    def productIds(self) -> list:
        return self._productIds

    @productIds.setter
    def productIds(self, ids: list):
        self._productIds = ids

    def compileModel(self, distributedConfig, numUser: int, numItem: int, numFactor: int = 10):
        spec = dict(SCRIPT_SPEC); spec["numFactor"] = numFactor
        self.model = NeuMFNet(numUser, numItem, dropout=0.0, seed=self.seed, **spec)
        return self.model

    def getPredictDataFrame(self, customerId):
        """NCFModel.py:32-35: one row per product id, the customer repeated."""
        return {'PRODUCT_ID': list(self.productIds), 'CUSTOMER_ID': [customerId] * len(self.productIds)}

    def checkpointMeta(self) -> dict:
        return {"model": self.modelName, "numUser": self.model.numUser, "numItem": self.model.numItem,
                "numFactor": self.model.E, "productIds": [int(x) for x in self.productIds],
                "customerIds": [int(x) for x in self.customerIds]}

    def buildFromMeta(self, meta: dict):
        self.compileModel(None, meta["numUser"], meta["numItem"], meta.get("numFactor", 10))
        if not self.productIds:
            self.productIds = list(meta.get("productIds", []))
        if not self.customerIds:
            self.customerIds = list(meta.get("customerIds", []))

    def predictForUsers(self, customerIds, numberOfItem=5):
        """[{product: '%.9f' % score, ...} per customer], best first (dicts keep insertion order)."""
        customerIds = [int(c) for c in customerIds]
        if self.model is None:
            raise RuntimeError("NCFModel: restoreFromLatestCheckPoint() or compileModel() first")
        bad = [c for c in customerIds if not 0 <= c < self.model.numUser]
        if bad:
            raise ValueError(f"unknown customer ids {bad[:5]}")
        if not customerIds or not self.productIds:
            return [{} for _ in customerIds]
        dev = self.model.device
        items = torch.as_tensor(np.asarray(self.productIds, dtype=np.int32)).to(dev)
        I = items.numel()
        k = min(int(numberOfItem), I)
        if not 1 <= k <= 32:
            raise ValueError("numberOfItem must be between 1 and 32 (the device top-k keeps at most 32 per row)")
        out = []
        per = max(1, (1 << 22) // I)                                      # <= 4 M (customer, product) pairs per pass
        for a in range(0, len(customerIds), per):
            chunk = customerIds[a:a + per]
            u = torch.as_tensor(np.asarray(chunk, dtype=np.int32)).to(dev).repeat_interleave(I)
            pred, _ = self.model.predict_on_batch(u, items.repeat(len(chunk)))
            scores = pred.view(len(chunk), I)
            scores = torch.where(scores < 1.0, scores, torch.full_like(scores, float("-inf")))   # filter(... < 1.0), :46
            v, ix = H.topk_rows(scores.contiguous(), k)
            v, ix = v.cpu().numpy(), ix.cpu().numpy()
            for r in range(len(chunk)):
                out.append({self.productIds[j]: '%.9f' % s for s, j in zip(v[r], ix[r]) if np.isfinite(s)})
        return out

    def predictForUser(self, customerId, numberOfItem=5):
        return self.predictForUsers([customerId], numberOfItem)[0]

    # This model only supports predict (NCFModel.py:53-55)
    def readyToTrain(self):
        return False

    def getPredictableUsers(self):
        return self.customerIds

    def loadTrainData(self):
        """customer_id,normalized_customer_id,material,product_id,rating_type (NCFModel.py:60-64): the distinct product
        ids and normalized customer ids of the file, in order of first appearance (`.unique()`)."""
        from .interactions import read_csv_columns
        with self.dataStore.openFile(self.trainData) as f:
            cols = read_csv_columns(f, "ncf")
        first = lambda a: list(dict.fromkeys(np.asarray(a).tolist()))
        self.productIds = first(cols["item"])
        self.customerIds = first(cols["user"])
