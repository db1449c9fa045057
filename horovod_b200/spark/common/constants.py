"""Names shared by the estimator pipeline (reference horovod/spark/common/constants.py)."""
# how a DataFrame column is laid out in the intermediate Parquet
ARRAY = 'array'                            # fixed-length list column, read back as an ndarray of that row shape
CUSTOM_SPARSE = 'custom_sparse_format'     # [size, nnz, indices..., values...] packing of a Spark SparseVector
NOCHANGE = 'nochange'                      # scalars and anything else written as is

# what kind of vector column the rows held before they were written
MIXED_SPARSE_DENSE_VECTOR = 'mixed_sparse_dense_vector'
SPARSE_VECTOR = 'sparse_vector'
DENSE_VECTOR = 'dense_vector'

METRIC_PRINT_FREQUENCY = 100               # steps between metric lines in verbose mode
TOTAL_BUFFER_MEMORY_CAP_GIB = 4            # cap for shuffle buffers sized from the average row size
BYTES_PER_GIB = 1 << 30
PETASTORM_HDFS_DRIVER = 'libhdfs'          # accepted for API parity; the reader here is pyarrow's own HDFS client
