"""Import-time stand-in for `soundfile` (absent offline). Test infrastructure only.
The reference imports it at distilcodec/distil_codec.py:12 and calls sf.write only in save_wav (:652)."""


def write(*a, **k):
    raise NotImplementedError("soundfile is not available in this environment")


def read(*a, **k):
    raise NotImplementedError("soundfile is not available in this environment")
