"""base64(cloudpickle) helpers used to smuggle objects through env vars / command lines."""
import base64

import cloudpickle


def dumps_base64(obj, to_ascii=True):
    serialized = base64.b64encode(cloudpickle.dumps(obj))
    return serialized.decode('ascii') if to_ascii else serialized


def loads_base64(encoded):
    return cloudpickle.loads(base64.b64decode(encoded))
