from _overlay import extend_path as _ext
__path__ = _ext(__path__, __name__)
