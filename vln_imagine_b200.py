"""Import shim: the package directory is ``vln-imagine_b200/`` (the hyphen is the project's name
and is not a legal Python identifier); ``import vln_imagine_b200`` resolves to that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'vln-imagine_b200')]
__file__ = _os.path.join(__path__[0], '__init__.py')
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, 'exec'))
