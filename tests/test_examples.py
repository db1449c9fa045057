"""The example programs run end to end on CPU (np=2) — they are the documentation users copy from."""
import os
import subprocess
import sys

from conftest import REPO


def _hvdrun(np_, *cmd, timeout=300):
    e = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', ''), HOROVOD_LOG_LEVEL='warning', OMP_NUM_THREADS='1')
    p = subprocess.run([sys.executable, '-m', 'horovod_b200.runner.launch', '-np', str(np_), sys.executable, *cmd],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=e, timeout=timeout, cwd=REPO)
    return p.returncode, p.stdout.decode(errors='replace')


def test_mnist_example(native_built):
    rc, out = _hvdrun(2, 'examples/pytorch_mnist_synthetic.py', '--epochs', '1', '--no-cuda')
    assert rc == 0, out[-3000:]


def test_imagenet_example_checkpoint_and_resume(native_built, tmp_path):
    fmt = str(tmp_path / 'ckpt-{epoch}.pt')
    common = ['examples/pytorch_imagenet_resnet50_synthetic.py', '--no-cuda', '--model', 'tiny', '--steps-per-epoch', '3', '--batch-size', '4',
              '--image-size', '32', '--checkpoint-format', fmt]
    rc, out = _hvdrun(2, *common, '--epochs', '2')
    assert rc == 0 and 'resumed from epoch 0' in out and os.path.exists(fmt.format(epoch=2)), out[-3000:]
    rc, out = _hvdrun(2, *common, '--epochs', '3', '--batches-per-allreduce', '2', '--compression', 'fp16')
    assert rc == 0 and 'resumed from epoch 2' in out and 'epoch 3:' in out and 'epoch 1:' not in out, out[-3000:]


def test_adasum_example(native_built):
    rc, out = _hvdrun(2, 'examples/adasum_small_model.py', '--no-cuda', '--steps', '60')
    assert rc == 0 and 'ADASUM EXAMPLE OK' in out, out[-3000:]


def test_numpy_example(native_built):
    rc, out = _hvdrun(3, 'examples/numpy_allreduce.py')
    assert rc == 0 and 'NUMPY EXAMPLE OK' in out, out[-3000:]


def _python(*cmd, timeout=600):
    e = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get('PYTHONPATH', ''), HOROVOD_LOG_LEVEL='warning', OMP_NUM_THREADS='1')
    p = subprocess.run([sys.executable, *cmd], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=e, timeout=timeout, cwd=REPO)
    return p.returncode, p.stdout.decode(errors='replace')


def test_estimator_examples(native_built, tmp_path):
    rc, out = _python('examples/spark_torch_estimator.py', '--num-proc', '2', '--epochs', '2', '--store', str(tmp_path / 's1'),
                      '--save-model', str(tmp_path / 'saved'))
    assert rc == 0 and 'ESTIMATOR EXAMPLE OK' in out and 'saved to and reloaded from' in out, out[-3000:]
    rc, out = _python('examples/spark_torch_estimator.py', '--num-proc', '2', '--epochs', '2', '--lightning', '--store', str(tmp_path / 's2'))
    assert rc == 0 and 'ESTIMATOR EXAMPLE OK' in out, out[-3000:]


def test_ray_executor_example(native_built):
    rc, out = _python('examples/ray_executor.py', '--num-workers', '2', '--steps', '60')
    assert rc == 0 and 'RAY EXAMPLE OK' in out, out[-3000:]
