"""tf.data service compute workers on Spark (reference horovod/spark/tensorflow/)."""
