"""Is a finished solver freed by reference counting alone (so that its HBM returns to the caching allocator at `del`)?"""
import gc, logging, os, sys, weakref, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import pymgrit_b200 as P
gc.disable()
nt, co = 2 ** 16 + 1, (16, 16)
s = P.Mgrit(problem=bench.hierarchy(P.Heat1D, nt, co), logging_lvl=logging.WARNING, **bench.SOLVER_KW)
s.solve()
w = weakref.ref(s)
lv = weakref.ref(s._lv[0])
u = weakref.ref(s._lv[0].u)
del s
print('solver alive after del:', w() is not None, '| level alive:', lv() is not None, '| u tensor alive:', u() is not None)
if w() is not None or lv() is not None:
    gc.collect()
    print('after gc.collect():', w() is not None, lv() is not None, u() is not None)
    gc.set_debug(gc.DEBUG_SAVEALL)
    s = P.Mgrit(problem=bench.hierarchy(P.Heat1D, nt, co), logging_lvl=logging.WARNING, **bench.SOLVER_KW)
    del s
    gc.collect()
    kinds = {}
    for o in gc.garbage:
        kinds[type(o).__name__] = kinds.get(type(o).__name__, 0) + 1
    print(sorted(kinds.items(), key=lambda kv: -kv[1])[:25])
    for o in gc.garbage:
        if type(o).__name__ in ('function', 'cell', 'method'):
            print(type(o).__name__, getattr(o, '__qualname__', ''), str(o)[:120])
