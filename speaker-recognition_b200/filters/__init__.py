from .silence import remove_silence  # noqa: F401
