"""Run artefact store for the estimators (parity: horovod/spark/common/store.py: `Store` interface :38-165,
`FilesystemStore`/`LocalStore` :301-395; HDFS/DBFS stores map onto `FilesystemStore` with an fsspec/pyarrow
filesystem).  Layout under `prefix_path`:

    intermediate_train_data[.idx]/   parquet written by the estimator
    intermediate_val_data[.idx]/
    runs/<run_id>/checkpoint.pt      rank-0 checkpoints
    runs/<run_id>/logs/
"""
import contextlib
import os
import shutil
import tempfile


def split_protocol(path):
    """'s3://bucket/x' -> ('s3', 'bucket/x'); 'C:/x' and '/x' -> (None, path) (reference store.py `split_protocol`)."""
    if '://' in path:
        scheme, rest = path.split('://', 1)
        if len(scheme) > 1:
            return scheme, rest
    return None, path


class _RemoteStore:
    """Plain, picklable snapshot of the paths and helpers a training process needs (what `Store.to_remote` returns)."""

    def __init__(self, attrs):
        self.__dict__.update(attrs)


class Store:
    """Interface; `Store.create(path)` picks an implementation from the path's scheme."""

    def get_train_data_path(self, idx=None):
        raise NotImplementedError

    def get_val_data_path(self, idx=None):
        raise NotImplementedError

    def get_test_data_path(self, idx=None):
        raise NotImplementedError

    def get_run_path(self, run_id):
        raise NotImplementedError

    def get_checkpoint_path(self, run_id):
        raise NotImplementedError

    def get_logs_path(self, run_id):
        raise NotImplementedError

    def exists(self, path):
        raise NotImplementedError

    def read(self, path):
        raise NotImplementedError

    def write(self, path, data):
        raise NotImplementedError

    def write_text(self, path, text):
        self.write(path, text.encode('utf-8'))

    def is_parquet_dataset(self, path):
        raise NotImplementedError

    def get_parquet_dataset(self, path):
        raise NotImplementedError

    def saving_runs(self):
        raise NotImplementedError

    def get_runs_path(self):
        raise NotImplementedError

    def get_checkpoints(self, run_id, suffix='.ckpt'):
        raise NotImplementedError

    def get_checkpoint_filename(self):
        raise NotImplementedError

    def get_logs_subdir(self):
        raise NotImplementedError

    def get_local_output_dir_fn(self, run_id):
        raise NotImplementedError

    def sync_fn(self, run_id):
        """fn(local_run_path): uploads what a training process wrote locally into the run's directory of the store."""
        raise NotImplementedError

    def to_remote(self, run_id, dataset_idx):
        """Snapshot for the training processes: attributes instead of methods, nothing that needs Spark or a live
        filesystem handle to unpickle (reference store.py:130-156)."""
        return _RemoteStore(self._remote_attrs(run_id, dataset_idx))

    def _remote_attrs(self, run_id, dataset_idx):
        saving = self.saving_runs()
        return {
            'train_data_path': self.get_train_data_path(dataset_idx),
            'val_data_path': self.get_val_data_path(dataset_idx),
            'test_data_path': self.get_test_data_path(dataset_idx),
            'saving_runs': saving,
            'runs_path': self.get_runs_path(),
            'run_path': self.get_run_path(run_id),
            'checkpoint_path': self.get_checkpoint_path(run_id),
            'logs_path': self.get_logs_path(run_id),
            'checkpoint_filename': self.get_checkpoint_filename(),
            'logs_subdir': self.get_logs_subdir(),
            'get_local_output_dir': self.get_local_output_dir_fn(run_id),
            'sync': self.sync_fn(run_id),
        }

    @staticmethod
    def create(prefix_path, *args, **kwargs):
        """hdfs://... -> HDFSStore; dbfs:/... or /dbfs/... -> DBFSLocalStore; other URIs -> FilesystemStore; plain paths -> LocalStore."""
        if DBFSLocalStore.matches(prefix_path):
            return DBFSLocalStore(prefix_path, *args, **kwargs)
        scheme = prefix_path.split('://', 1)[0] if '://' in prefix_path else 'file'
        if scheme == 'file':
            return LocalStore(prefix_path, *args, **kwargs)
        if scheme in ('hdfs', 'viewfs'):
            return HDFSStore(prefix_path, *args, **kwargs)
        return FilesystemStore(prefix_path, *args, **kwargs)


class FilesystemStore(Store):
    """Any pyarrow.fs filesystem (local, hdfs://, s3://, gs://); paths handed out are URIs usable by pyarrow.dataset."""

    def __init__(self, prefix_path, train_path=None, val_path=None, test_path=None, runs_path=None, save_runs=True, filesystem=None):
        self.prefix_path = prefix_path.rstrip('/')
        self._train = train_path or self._join('intermediate_train_data')
        self._val = val_path or self._join('intermediate_val_data')
        self._test = test_path or self._join('intermediate_test_data')
        self._runs = runs_path or self._join('runs')
        self._save_runs = save_runs
        self._fs = filesystem

    def _join(self, *parts):
        return '/'.join([self.prefix_path] + list(parts))

    @property
    def fs(self):
        if self._fs is None:
            import pyarrow.fs as pafs
            self._fs, self._root = pafs.FileSystem.from_uri(self.prefix_path)
        return self._fs

    def _local(self, path):
        """Path as the filesystem object wants it (URI scheme stripped)."""
        return path.split('://', 1)[1] if '://' in path else path

    def get_train_data_path(self, idx=None):
        return self._train if idx is None else f'{self._train}.{idx}'

    def get_val_data_path(self, idx=None):
        return self._val if idx is None else f'{self._val}.{idx}'

    def get_test_data_path(self, idx=None):
        return self._test if idx is None else f'{self._test}.{idx}'

    def saving_runs(self):
        return self._save_runs

    def get_runs_path(self):
        return self._runs

    def get_run_path(self, run_id):
        return f'{self._runs}/{run_id}'

    def get_checkpoint_filename(self):
        return 'checkpoint.pt'

    def get_checkpoint_path(self, run_id):
        return f'{self.get_run_path(run_id)}/{self.get_checkpoint_filename()}' if self._save_runs else None

    def get_logs_subdir(self):
        return 'logs'

    def get_logs_path(self, run_id):
        return f'{self.get_run_path(run_id)}/{self.get_logs_subdir()}' if self._save_runs else None

    def exists(self, path):
        import pyarrow.fs as pafs
        return self.fs.get_file_info(self._local(path)).type != pafs.FileType.NotFound

    def read(self, path):
        with self.fs.open_input_stream(self._local(path)) as f:
            return f.read()

    def write(self, path, data):
        p = self._local(path)
        self.fs.create_dir(os.path.dirname(p), recursive=True)
        with self.fs.open_output_stream(p) as f:
            f.write(data)

    def is_parquet_dataset(self, path):
        import pyarrow.fs as pafs
        info = self.fs.get_file_info(pafs.FileSelector(self._local(path), allow_not_found=True))
        return any(i.path.endswith('.parquet') for i in info)

    def get_parquet_dataset(self, path):
        import pyarrow.parquet as pq
        return pq.ParquetDataset(self._local(path), filesystem=self.fs)

    def get_data_metadata_path(self, path):
        """Where the estimator keeps the row count / schema summary of a materialised DataFrame."""
        return self.get_localized_path(path).rstrip('/') + '/_metadata.json'

    def get_checkpoints(self, run_id, suffix='.ckpt'):
        """Files below the run directory that end in `suffix` (what the Lightning-style trainers write per epoch)."""
        import pyarrow.fs as pafs
        root = self._local(self.get_run_path(run_id))
        info = self.fs.get_file_info(pafs.FileSelector(root, recursive=True, allow_not_found=True))
        return sorted(i.path for i in info if i.type == pafs.FileType.File and i.path.endswith(suffix))

    def get_localized_path(self, path):
        """`path` as this store's filesystem object addresses it (scheme and authority removed)."""
        return self._local(path)

    def get_full_path(self, path):
        """`path` as a URI another process can open without this object."""
        if split_protocol(path)[0] is not None:
            return path
        scheme = split_protocol(self.prefix_path)[0]
        return path if scheme is None else f'{scheme}://{path.lstrip("/") if scheme not in ("file", "hdfs", "viewfs") else path}'

    def get_full_path_fn(self):
        prefix = split_protocol(self.prefix_path)[0]

        def full(path):
            return path if prefix is None or '://' in path else f'{prefix}://{path}'
        return full

    def get_filesystem(self):
        return self.fs

    @classmethod
    def matches(cls, path):
        """Any URI with a scheme (the most general store; `Store.create` tries the specific ones first)."""
        return split_protocol(path)[0] is not None

    def get_local_output_dir_fn(self, run_id):
        """Training processes write checkpoints / logs into a scratch directory first; `sync_fn` uploads it."""
        @contextlib.contextmanager
        def local_run_path():
            d = tempfile.mkdtemp(prefix='hvd_run_')
            try:
                yield d
            finally:
                shutil.rmtree(d, ignore_errors=True)
        return local_run_path

    def copy(self, lpath, rpath, recursive=False):
        """Local file or directory -> this store."""
        import pyarrow.fs as pafs
        dst = self._local(rpath)
        if os.path.isdir(lpath):
            if not recursive:
                raise IsADirectoryError(lpath)
            pafs.copy_files(lpath, dst, source_filesystem=pafs.LocalFileSystem(), destination_filesystem=self.fs)
        else:
            self.fs.create_dir(os.path.dirname(dst), recursive=True)
            with open(lpath, 'rb') as f, self.fs.open_output_stream(dst) as out:
                shutil.copyfileobj(f, out)

    def sync_fn(self, run_id):
        run_path, copy = self.get_run_path(run_id), self.copy

        def fn(local_run_path):
            copy(local_run_path, run_path, recursive=True)
        return fn

    def read_serialized_keras_model(self, ckpt_path, model, custom_objects):
        """The checkpoint written by the Keras estimator, re-encoded the way `KerasModel` keeps models (bytes of an
        in-memory h5 / weights blob).  `model` supplies the architecture when the checkpoint holds weights only."""
        from horovod_b200.spark.keras import util as kutil
        return kutil.checkpoint_to_serialized_model(self.read(ckpt_path), model, custom_objects)

    def delete(self, path):
        if self.exists(path):
            self.fs.delete_dir(self._local(path))


class LocalStore(FilesystemStore):
    """Plain directories on a filesystem every worker can see (single box, NFS)."""

    def __init__(self, prefix_path, *args, **kwargs):
        if prefix_path.startswith('file://'):
            prefix_path = prefix_path[len('file://'):]
        super().__init__(os.path.abspath(prefix_path), *args, **kwargs)

    @property
    def fs(self):
        if self._fs is None:
            import pyarrow.fs as pafs
            self._fs = pafs.LocalFileSystem()
        return self._fs

    def delete(self, path):
        shutil.rmtree(path, ignore_errors=True)

    @classmethod
    def matches(cls, path):
        return split_protocol(path)[0] in (None, 'file')

    def get_full_path(self, path):
        return path if '://' in path else 'file://' + os.path.abspath(path)

    def get_full_path_fn(self):
        return lambda path: path if '://' in path else 'file://' + os.path.abspath(path)


class HDFSStore(FilesystemStore):
    """hdfs://[host[:port]]/path (reference store.py `HDFSStore`): a FilesystemStore on pyarrow's HadoopFileSystem.  Host / port /
    user / kerberos ticket can be given explicitly; otherwise they come from the URI (or from the Hadoop configuration when the
    URI has no authority: `hdfs:///path`)."""

    def __init__(self, prefix_path, host=None, port=None, user=None, kerb_ticket=None, extra_conf=None, *args, **kwargs):
        self._hdfs_kwargs = dict(host=host, port=port, user=user, kerb_ticket=kerb_ticket, extra_conf=extra_conf)
        super().__init__(prefix_path, *args, **kwargs)

    @staticmethod
    def parse_url(url):
        """-> (host or 'default', port or 0, path)"""
        rest = url.split('://', 1)[1] if '://' in url else url
        authority, _, path = rest.partition('/')
        host, _, port = authority.partition(':')
        return host or 'default', int(port) if port else 0, '/' + path

    @property
    def fs(self):
        if self._fs is None:
            import pyarrow.fs as pafs
            host, port, _ = self.parse_url(self.prefix_path)
            kw = {k: v for k, v in self._hdfs_kwargs.items() if v is not None}
            kw.setdefault('host', host)
            kw.setdefault('port', port)
            self._fs = pafs.HadoopFileSystem(**kw)
        return self._fs

    def _local(self, path):
        return self.parse_url(path)[2] if '://' in path else path

    @classmethod
    def matches(cls, path):
        return split_protocol(path)[0] in ('hdfs', 'viewfs')

    def get_full_path(self, path):
        if '://' in path:
            return path
        scheme, rest = split_protocol(self.prefix_path)
        return f'{scheme}://{rest.partition("/")[0]}{path}'

    def get_full_path_fn(self):
        scheme, rest = split_protocol(self.prefix_path)
        authority = rest.partition('/')[0]
        return lambda path: path if '://' in path else f'{scheme}://{authority}{path}'


def is_databricks():
    return 'DATABRICKS_RUNTIME_VERSION' in os.environ


class DBFSLocalStore(LocalStore):
    """Databricks file system through its FUSE mount: `dbfs:/x` and `/dbfs/x` both mean the local path `/dbfs/x` (reference
    store.py `DBFSLocalStore`).  Checkpoints are written with the `.tf` suffix the reference uses there."""

    def __init__(self, prefix_path, *args, **kwargs):
        super().__init__(self.normalize_path(prefix_path), *args, **kwargs)

    @staticmethod
    def matches(path):
        return path.startswith('dbfs:/') or path == '/dbfs' or path.startswith('/dbfs/')

    matches_dbfs = matches

    def get_localized_path(self, path):
        return self.normalize_path(path)

    def get_full_path(self, path):
        return 'file://' + self.normalize_path(path)

    @staticmethod
    def normalize_path(path):
        if path.startswith('dbfs:///'):
            return '/dbfs/' + path[len('dbfs:///'):]
        if path.startswith('dbfs:/'):
            return '/dbfs/' + path[len('dbfs:/'):].lstrip('/')
        return path

    def get_checkpoint_filename(self):
        return 'checkpoint.tf'


AbstractFilesystemStore = FilesystemStore
