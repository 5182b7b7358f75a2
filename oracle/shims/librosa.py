"""Import-time stand-in for `librosa` (absent offline). Test infrastructure only.
Reference import sites: distilcodec/distil_codec.py:14, distilcodec/models/meldataset.py:9."""


def load(*a, **k):
    raise NotImplementedError("librosa is not available in this environment")


def resample(*a, **k):
    raise NotImplementedError("librosa is not available in this environment")


class util:  # meldataset.py touches librosa.util at import in some versions
    @staticmethod
    def normalize(x, *a, **k):
        raise NotImplementedError
