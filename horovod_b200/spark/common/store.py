"""Where estimators keep intermediate data, checkpoints and logs (role parity: horovod/spark/common/store.py:38-165).
LocalStore works on any mounted filesystem; HDFS/DBFS variants need their client libraries and are gated."""
import os
import shutil


class Store(object):
    """Abstracts reading and writing of intermediate data and run results."""

    def is_parquet_dataset(self, path):
        raise NotImplementedError()

    def get_train_data_path(self, idx=None):
        raise NotImplementedError()

    def get_val_data_path(self, idx=None):
        raise NotImplementedError()

    def get_test_data_path(self, idx=None):
        raise NotImplementedError()

    def saving_runs(self):
        raise NotImplementedError()

    def get_runs_path(self):
        raise NotImplementedError()

    def get_run_path(self, run_id):
        raise NotImplementedError()

    def get_checkpoint_path(self, run_id):
        raise NotImplementedError()

    def get_logs_path(self, run_id):
        raise NotImplementedError()

    def get_checkpoint_filename(self):
        raise NotImplementedError()

    def get_logs_subdir(self):
        raise NotImplementedError()

    def exists(self, path):
        raise NotImplementedError()

    def read(self, path):
        raise NotImplementedError()

    def write(self, path, data):
        raise NotImplementedError()

    @staticmethod
    def create(prefix_path, *args, **kwargs):
        if prefix_path.startswith('hdfs://'):
            raise ImportError('HDFSStore needs pyarrow.hdfs / fsspec[hdfs], not available in this environment')
        if prefix_path.startswith('dbfs:/'):
            raise ImportError('DBFSLocalStore is only meaningful on Databricks')
        return LocalStore(prefix_path, *args, **kwargs)


class FilesystemStore(Store):
    """Store on a POSIX-like filesystem layout: <prefix>/intermediate_{train,val,test}_data, <prefix>/runs/<id>/..."""

    def __init__(self, prefix_path, train_path=None, val_path=None, test_path=None, runs_path=None, save_runs=True):
        self.prefix_path = self.get_full_path(prefix_path)
        self._train_path = self._get_full_path_or_default(train_path, 'intermediate_train_data')
        self._val_path = self._get_full_path_or_default(val_path, 'intermediate_val_data')
        self._test_path = self._get_full_path_or_default(test_path, 'intermediate_test_data')
        self._runs_path = self._get_full_path_or_default(runs_path, 'runs')
        self._save_runs = save_runs

    def _get_full_path_or_default(self, path, default_key):
        return self.get_full_path(path) if path is not None else self._get_path(default_key)

    def _get_path(self, key):
        return os.path.join(self.prefix_path, key)

    def get_full_path(self, path):
        return os.path.abspath(path)

    def _indexed(self, base, idx):
        return '{}.{}'.format(base, idx) if idx is not None else base

    def get_train_data_path(self, idx=None):
        return self._indexed(self._train_path, idx)

    def get_val_data_path(self, idx=None):
        return self._indexed(self._val_path, idx)

    def get_test_data_path(self, idx=None):
        return self._indexed(self._test_path, idx)

    def is_parquet_dataset(self, path):
        return os.path.isdir(path) and any(f.endswith('.parquet') for f in os.listdir(path))

    def saving_runs(self):
        return self._save_runs

    def get_runs_path(self):
        return self._runs_path

    def get_run_path(self, run_id):
        return os.path.join(self.get_runs_path(), run_id)

    def get_checkpoint_path(self, run_id):
        return os.path.join(self.get_run_path(run_id), self.get_checkpoint_filename()) if self._save_runs else None

    def get_logs_path(self, run_id):
        return os.path.join(self.get_run_path(run_id), self.get_logs_subdir()) if self._save_runs else None

    def get_checkpoint_filename(self):
        return 'checkpoint.pt'

    def get_logs_subdir(self):
        return 'logs'

    def exists(self, path):
        return os.path.exists(path)

    def read(self, path):
        with open(path, 'rb') as f:
            return f.read()

    def write(self, path, data):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + '.tmp'
        with open(tmp, 'wb') as f:
            f.write(data)
        os.replace(tmp, path)  # atomic: a reader never sees a half-written checkpoint

    def sync_fn(self, run_id):
        run_path = self.get_run_path(run_id)

        def fn(local_run_path):
            if os.path.abspath(local_run_path) != os.path.abspath(run_path):
                shutil.copytree(local_run_path, run_path, dirs_exist_ok=True)
        return fn


class LocalStore(FilesystemStore):
    """Uses the local filesystem as a store of intermediate data and training artifacts."""
    FS_PREFIX = 'file://'

    def __init__(self, prefix_path, *args, **kwargs):
        if prefix_path.startswith(self.FS_PREFIX):
            prefix_path = prefix_path[len(self.FS_PREFIX):]
        super().__init__(prefix_path, *args, **kwargs)
