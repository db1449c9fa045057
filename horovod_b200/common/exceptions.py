"""Exception types surfaced to user code (reference horovod/common/exceptions.py)."""


class HorovodInternalError(RuntimeError):
    """Internal error raised when a collective fails (peer death, mismatched submission, shutdown).

    In elastic mode it triggers restore-from-last-commit + re-initialisation."""


class HostsUpdatedInterrupt(RuntimeError):
    """Raised inside `state.commit()` / `state.check_host_updates()` when the driver announced a host change.

    `skip_sync` is True when the change only removed hosts (no state re-broadcast needed)."""

    def __init__(self, skip_sync=False):
        super().__init__()
        self.skip_sync = skip_sync


class HorovodVersionMismatchError(ImportError):
    def __init__(self, name, version, installed_version):
        super().__init__(f"Framework {name} installed with version {installed_version} but found version {version}.")
        self.name = name
        self.version = version
        self.installed_version = installed_version


def get_version_mismatch_message(name, version, installed_version):
    return (f"Framework {name} installed with version {installed_version} but found version {version}.\n"
            f"             This can result in unexpected behavior including runtime errors.\n"
            f"             Rebuild the native library: python -m horovod_b200.build --force")
