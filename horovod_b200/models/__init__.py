"""Benchmark model zoo (synthetic-data drivers only: the library itself is model-agnostic, like the reference, whose
only models live in examples/).  ResNet-50, BERT-large and GPT-2-medium are the configs named by BASELINE.json."""
from horovod_b200.models.resnet import resnet50, resnet101  # noqa: F401
from horovod_b200.models.bert import BertConfig, BertForPreTraining, bert_large  # noqa: F401
from horovod_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel, gpt2_medium  # noqa: F401
