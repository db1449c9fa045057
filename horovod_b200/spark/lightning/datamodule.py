"""Data module of the Lightning estimator (reference horovod/spark/lightning/datamodule.py `PetastormDataModule` :17-160: a
LightningDataModule over Petastorm readers with optional async loaders).  Same reader as the torch estimator; the extra
methods are the LightningDataModule names, so the object can be handed to code that expects one."""
from horovod_b200.spark.torch.datamodule import MapIterable, ParquetDataModule as _TorchParquetDataModule  # noqa: F401


class ParquetDataModule(_TorchParquetDataModule):
    def setup(self, stage=None):
        return self.__enter__()

    def teardown(self, stage=None):
        return self.__exit__(None, None, None)

    def train_dataloader(self):
        return self.train_data()

    def val_dataloader(self):
        return self.val_data()


PetastormDataModule = ParquetDataModule
