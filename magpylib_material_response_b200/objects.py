"""Minimal duck-typed stand-ins for the magpylib objects the demag path touches.

magpylib (the reference's object model, /root/reference/src/magpylib_material_response/
demag.py:7-11, meshing.py:6-9) is not installable in this environment, so ``Cuboid``,
``Collection`` and ``Sensor`` are re-stated here with exactly the attributes the hot path
reads: ``position``, ``orientation`` (scipy Rotation), ``dimension``, ``polarization``,
``barycenter``, ``parent``, ``children``/``sources_all``, ``copy()``, ``style.label`` and
``getB``/``getH``.  When magpylib IS importable, ``_compat`` hands out the real classes
instead and this module is unused.

Field evaluation (``getB``/``getH``) goes to the CUDA field kernel through the C ABI
(``mmr_field``); there is no numpy implementation of the field in the product path.
"""

from __future__ import annotations

import copy as _copy

import numpy as np
from scipy.spatial.transform import Rotation as R

MU0 = 1.25663706127e-06  # = scipy.constants.mu_0 = magpylib.mu_0


class Style:
    """Attribute bag with magpylib-like ``update`` / ``as_dict`` (display only)."""

    def __init__(self, **kwargs):
        self.label = None
        self.opacity = None
        self.color = None
        self.description = None
        self.update(kwargs)

    def update(self, arg=None, _match_properties=True, **kwargs):
        items = dict(arg or {})
        items.update(kwargs)
        for key, val in items.items():
            if val is None:
                continue
            if isinstance(val, dict) and isinstance(getattr(self, key, None), Style):
                getattr(self, key).update(val)
            else:
                setattr(self, key, val)
        return self

    def as_dict(self):
        return {k: (v.as_dict() if isinstance(v, Style) else v) for k, v in vars(self).items()}

    def copy(self):
        new = Style.__new__(Style)
        for k, v in vars(self).items():
            new.__dict__[k] = v.copy() if isinstance(v, Style) else (_copy.deepcopy(v) if isinstance(v, dict | list) else v)
        return new

    def __repr__(self):
        return f"Style({self.as_dict()})"


_IDENTITY = R.identity()  # shared default (Rotation objects are immutable)


def _as_rotation(orientation):
    if orientation is None:
        return _IDENTITY
    if isinstance(orientation, R):
        return orientation
    msg = "orientation must be None or a scipy.spatial.transform.Rotation"
    raise TypeError(msg)


def _rot_single(rot):
    """A Rotation holding exactly one rotation as a 'single' object."""
    q = np.asarray(rot.as_quat())
    if q.ndim == 2 and len(q) == 1:
        return R.from_quat(q[0])
    return rot


class BaseGeo:
    """position / orientation / parent / style (+ move, rotate) of a magpylib object."""

    _ids = 0

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, style=None, **kwargs):
        self._position = np.array(position, dtype=float)
        self._orientation = _rot_single(_as_rotation(orientation))
        self._parent = None
        style_kwargs = {k[6:]: v for k, v in kwargs.items() if k.startswith("style_")}
        other = [k for k in kwargs if not k.startswith("style_")]
        if other:
            msg = f"{type(self).__name__}() got unexpected keyword arguments {other}"
            raise TypeError(msg)
        self.style = Style(**(style or {}))
        self.style.update(style_kwargs)
        BaseGeo._ids += 1
        self._id = BaseGeo._ids

    # -- geometry ------------------------------------------------------------------
    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, value):
        self._set_position(np.array(value, dtype=float))

    def _set_position(self, value):
        self._position = value

    @property
    def orientation(self):
        return self._orientation

    @orientation.setter
    def orientation(self, value):
        self._set_orientation(_rot_single(_as_rotation(value)))

    def _set_orientation(self, value):
        self._orientation = value

    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, value):
        if value is None:
            if self._parent is not None:
                self._parent.remove(self)
            self._parent = None
        else:
            value.add(self)

    def move(self, displacement, start="auto"):
        disp = np.array(displacement, dtype=float)
        if disp.ndim == 1 and self._position.ndim == 1:
            self.position = self._position + disp
        else:
            # path semantics (start=0 style): broadcast the object's pose along the path
            base = np.atleast_2d(self._position)
            disp = np.atleast_2d(disp)
            p = max(len(base), len(disp))
            base = np.vstack([base, np.repeat(base[-1:], p - len(base), axis=0)])
            disp = np.vstack([disp, np.repeat(disp[-1:], p - len(disp), axis=0)])
            self._pad_orientation(p)
            self.position = base + disp
        return self

    def _pad_orientation(self, p):
        q = np.atleast_2d(self._orientation.as_quat())
        if len(q) < p:
            q = np.vstack([q, np.repeat(q[-1:], p - len(q), axis=0)])
            self._orientation = R.from_quat(q)

    def rotate(self, rotation, anchor=None, start="auto"):
        rot = _as_rotation(rotation)
        rq = np.atleast_2d(rot.as_quat())
        pos = np.atleast_2d(self._position)
        oq = np.atleast_2d(self._orientation.as_quat())
        p = max(len(rq), len(pos), len(oq))
        single = p == 1 and self._position.ndim == 1 and np.asarray(rot.as_quat()).ndim == 1

        def pad(a):
            return np.vstack([a, np.repeat(a[-1:], p - len(a), axis=0)])

        rq, pos, oq = pad(rq), pad(pos), pad(oq)
        rot_p = R.from_quat(rq)
        new_o = rot_p * R.from_quat(oq)
        if anchor is not None:
            anc = np.array(anchor, dtype=float) * np.ones(3)
            pos = anc + rot_p.apply(pos - anc)
        if single:
            self._orientation = _rot_single(new_o)
            self.position = pos[0]
        else:
            self._orientation = new_o
            self.position = pos
        return self

    def rotate_from_angax(self, angle, axis, anchor=None, start="auto", degrees=True):
        if isinstance(axis, str):
            axis = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[axis]
        ax = np.array(axis, dtype=float)
        ax = ax / np.linalg.norm(ax)
        ang = np.array(angle, dtype=float)
        if degrees:
            ang = np.deg2rad(ang)
        rotvec = ax * ang if ang.ndim == 0 else ang[:, None] * ax[None, :]
        return self.rotate(R.from_rotvec(rotvec), anchor=anchor, start=start)

    def rotate_from_rotvec(self, rotvec, anchor=None, start="auto", degrees=True):
        rv = np.array(rotvec, dtype=float)
        if degrees:
            rv = np.deg2rad(rv)
        return self.rotate(R.from_rotvec(rv), anchor=anchor, start=start)

    # -- misc ----------------------------------------------------------------------
    def copy(self, **kwargs):
        new = _copy.deepcopy(self)
        new._parent = None
        for k, v in kwargs.items():
            setattr(new, k, v)
        return new

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_parent":
                # keep parent links only inside the copied subtree
                new._parent = memo.get(id(v)) if v is not None else None
            elif k == "_orientation":
                new._orientation = v  # Rotation objects are immutable
            elif isinstance(v, np.ndarray):
                new.__dict__[k] = v.copy()
            elif v is None or isinstance(v, int | float | str | bool | tuple):
                new.__dict__[k] = v  # immutable (tuples of numbers: susceptibility, H_ext)
            elif isinstance(v, Style):
                new.__dict__[k] = v.copy()
            else:
                setattr(new, k, _copy.deepcopy(v, memo))
        BaseGeo._ids += 1
        new._id = BaseGeo._ids
        return new

    def __repr__(self):
        lbl = self.style.label
        return f"{type(self).__name__}(id={self._id}" + (f", label={lbl!r})" if lbl else ")")


class BaseSource(BaseGeo):
    def getB(self, *observers, **kwargs):
        from .field import getBH

        return getBH("B", [self], _merge_observers(observers), **kwargs)

    def getH(self, *observers, **kwargs):
        from .field import getBH

        return getBH("H", [self], _merge_observers(observers), **kwargs)


class BaseMagnet(BaseSource):
    """Magnet base: polarization J (tesla) in the object's local frame."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, polarization=None,
                 magnetization=None, style=None, **kwargs):
        super().__init__(position, orientation, style, **kwargs)
        if polarization is not None and magnetization is not None:
            msg = "give polarization or magnetization, not both"
            raise ValueError(msg)
        self._polarization = None
        if magnetization is not None:
            self.magnetization = magnetization
        elif polarization is not None:
            self.polarization = polarization

    @property
    def polarization(self):
        return self._polarization

    @polarization.setter
    def polarization(self, value):
        if value is None:
            self._polarization = None
            return
        v = np.array(value, dtype=float)
        if v.shape != (3,):
            msg = f"polarization must be a 3-vector, got shape {v.shape}"
            raise ValueError(msg)
        self._polarization = v

    @property
    def magnetization(self):
        return None if self._polarization is None else self._polarization / MU0

    @magnetization.setter
    def magnetization(self, value):
        self.polarization = None if value is None else np.array(value, dtype=float) * MU0


class BaseCurrent(BaseSource):
    """Base of the line-current sources (magpylib.current.*)."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, current=None, style=None, **kwargs):
        super().__init__(position, orientation, style, **kwargs)
        self.current = current


class Polyline(BaseCurrent):
    """magpylib.current.Polyline stand-in: straight segments through ``vertices`` (local frame, m)."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, vertices=None, current=None, style=None,
                 **kwargs):
        super().__init__(position, orientation, current, style, **kwargs)
        self.vertices = vertices

    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, value):
        if value is None:
            self._vertices = None
            return
        arr = np.array(value, dtype=float)
        if arr.ndim != 2 or arr.shape[1] != 3 or arr.shape[0] < 2:
            msg = f"vertices must have shape (n>=2, 3), got {arr.shape}"
            raise ValueError(msg)
        self._vertices = arr


class Circle(BaseCurrent):
    """magpylib.current.Circle stand-in: loop of ``diameter`` (m) in the local xy plane."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, diameter=None, current=None, style=None,
                 **kwargs):
        super().__init__(position, orientation, current, style, **kwargs)
        self.diameter = None if diameter is None else float(diameter)


class Cuboid(BaseMagnet):
    """Homogeneously polarised cuboid, edge lengths ``dimension`` (metres)."""

    def __init__(self, position=(0.0, 0.0, 0.0), orientation=None, dimension=None, polarization=None,
                 magnetization=None, style=None, **kwargs):
        super().__init__(position, orientation, polarization, magnetization, style, **kwargs)
        self._dimension = None
        if dimension is not None:
            self.dimension = dimension

    @property
    def dimension(self):
        return self._dimension

    @dimension.setter
    def dimension(self, value):
        if value is None:
            self._dimension = None
            return
        v = np.array(value, dtype=float)
        if v.shape != (3,) or np.any(v <= 0):
            msg = f"dimension must be a positive 3-vector, got {value!r}"
            raise ValueError(msg)
        self._dimension = v

    @property
    def barycenter(self):
        return self._position

    @property
    def volume(self):
        return float(np.prod(self._dimension))


class CellBlock:
    """Structure-of-arrays storage of the n Cuboid cells of one meshed collection: ``pos`` (n,3) centres,
    ``dim`` (n,3) edge lengths, ``pol`` (n,3) local-frame polarizations with ``pol_set`` (n,) marking the
    cells whose polarization is not None, ONE orientation ``rot`` shared by all cells (what
    ``mesh_Cuboid`` produces: /root/reference/src/magpylib_material_response/meshing.py:60-95) and the
    attributes every cell carries alike (``attrs``: susceptibility, H_ext; meshing.py:19-29).

    The Python objects of the cells are ``CellView`` rows onto these arrays, created only when somebody
    asks for them (``collection.children``, ``sources_all`` ...); ``apply_demag`` / ``getB`` read and write
    the arrays directly, so the per-cell object cost of the reference (meshing.py:90-93, demag.py:380-381,
    :420-428, :472-476) disappears from the hot path.
    """

    __slots__ = ("pos", "dim", "pol", "pol_set", "rot", "attrs", "custom")

    def __init__(self, pos, dim, pol=None, rot=None, attrs=None):
        self.pos = np.ascontiguousarray(pos, dtype=float).reshape(-1, 3).copy()
        n = len(self.pos)
        self.dim = np.ascontiguousarray(np.broadcast_to(np.asarray(dim, dtype=float), (n, 3))).copy()
        self.pol = np.zeros((n, 3))
        self.pol_set = np.zeros(n, dtype=bool)
        if pol is not None:
            self.pol[:] = np.asarray(pol, dtype=float)
            self.pol_set[:] = True
        self.rot = _IDENTITY if rot is None else _rot_single(rot)
        self.attrs = dict(attrs or {})
        self.custom = False  # a cell view received an attribute of its own: array fast paths are off

    def __len__(self):
        return len(self.pos)

    def copy(self):
        new = CellBlock.__new__(CellBlock)
        new.pos, new.dim, new.pol, new.pol_set = self.pos.copy(), self.dim.copy(), self.pol.copy(), self.pol_set.copy()
        new.rot = self.rot
        new.attrs = _copy.deepcopy(self.attrs)
        new.custom = False
        return new

    def __deepcopy__(self, memo):
        return self.copy()

    def shift(self, delta):
        self.pos += delta

    def rotate_about(self, rel, anchor):
        self.pos[:] = anchor + rel.apply(self.pos - anchor)
        self.rot = _rot_single(rel * self.rot)


def _row_property(field):
    def fget(self):
        return getattr(self._blk, field)[self._i]

    return fget


class CellView(Cuboid):
    """A Cuboid whose position / dimension / polarization are rows of its collection's ``CellBlock``
    (reads return views, assignments write into the arrays).  Anything a block cannot represent -- a
    path, an orientation of its own, leaving the collection -- first turns every cell of the block into
    an ordinary ``Cuboid`` (``Collection._detach_block``)."""

    _OWN = frozenset(("_blk", "_i", "_parent", "_style", "_id"))

    def __init__(self, *args, **kwargs):
        msg = "CellView objects are created by their Collection"
        raise TypeError(msg)

    @classmethod
    def _make(cls, blk, i, parent):
        new = cls.__new__(cls)
        d = new.__dict__
        d["_blk"], d["_i"], d["_parent"], d["_style"] = blk, i, parent, None
        BaseGeo._ids += 1
        d["_id"] = BaseGeo._ids
        return new

    def _detached(self):
        """Detach the whole block; afterwards ``self`` is a plain Cuboid."""
        self.__dict__["_parent"]._detach_block()

    # pose --------------------------------------------------------------------------------------
    @property
    def _position(self):
        return self._blk.pos[self._i]

    @_position.setter
    def _position(self, value):
        value = np.asarray(value, dtype=float)
        if value.shape == (3,):
            self._blk.pos[self._i] = value
        else:
            self._detached()
            self._position = value

    @property
    def _orientation(self):
        return self._blk.rot

    @_orientation.setter
    def _orientation(self, value):
        if value is self._blk.rot:
            return
        self._detached()
        self._orientation = value

    @property
    def _dimension(self):
        return self._blk.dim[self._i]

    @_dimension.setter
    def _dimension(self, value):
        if value is None:
            self._detached()
            self._dimension = None
        else:
            self._blk.dim[self._i] = value

    @property
    def _polarization(self):
        blk, i = self._blk, self._i
        return blk.pol[i] if blk.pol_set[i] else None

    @_polarization.setter
    def _polarization(self, value):
        blk, i = self._blk, self._i
        if value is None:
            blk.pol[i] = 0.0
            blk.pol_set[i] = False
        else:
            blk.pol[i] = value
            blk.pol_set[i] = True

    @property
    def style(self):
        d = self.__dict__
        if d["_style"] is None:
            d["_style"] = Style()
        return d["_style"]

    @style.setter
    def style(self, value):
        self.__dict__["_style"] = value

    # attributes shared by all cells of the block (susceptibility, H_ext) ----------------------------
    def __getattr__(self, name):
        attrs = self.__dict__["_blk"].attrs if "_blk" in self.__dict__ else {}
        if name in attrs:
            return attrs[name]
        msg = f"{type(self).__name__!r} object has no attribute {name!r}"
        raise AttributeError(msg)

    def __setattr__(self, name, value):
        if name not in CellView._OWN and not isinstance(getattr(type(self), name, None), property):
            self.__dict__["_blk"].custom = True
        object.__setattr__(self, name, value)

    def __delattr__(self, name):
        if name not in self.__dict__ and name in self._blk.attrs:
            self._detached()
        object.__delattr__(self, name)

    def __deepcopy__(self, memo):
        # a cell copied on its own becomes an ordinary Cuboid
        new = Cuboid(position=self._position.copy(), orientation=self._orientation, dimension=self._dimension.copy(),
                     polarization=None if self._polarization is None else self._polarization.copy())
        new.style = self.style.copy()
        for k, v in {**self._blk.attrs, **{k: v for k, v in self.__dict__.items() if k not in CellView._OWN}}.items():
            setattr(new, k, _copy.deepcopy(v, memo))
        return new


class Sensor(BaseGeo):
    """Observer positions ``pixel`` in the sensor's local frame."""

    def __init__(self, position=(0.0, 0.0, 0.0), pixel=(0.0, 0.0, 0.0), orientation=None, style=None,
                 **kwargs):
        super().__init__(position, orientation, style, **kwargs)
        self.pixel = np.array(pixel, dtype=float)

    def global_pixels(self):
        pix = self.pixel.reshape(-1, 3)
        return self._orientation.apply(pix) + self._position


class Collection(BaseSource):
    """Group of objects; moving/rotating the collection moves/rotates its children.

    A collection produced by ``mesh_Cuboid`` keeps its cells in a ``CellBlock`` (structure of arrays); the
    ``CellView`` children are materialised on first access and every pose operation on the collection is
    an array operation on the block.  Changing the child list of such a collection (``add`` / ``remove``)
    converts it to ordinary ``Cuboid`` children first.
    """

    def __init__(self, *children, position=(0.0, 0.0, 0.0), orientation=None, style=None, **kwargs):
        super().__init__(position, orientation, style, **kwargs)
        self._block = None
        self._kids = []
        self.add(*children)

    @classmethod
    def _from_block(cls, block):
        new = cls()
        new._block = block
        new._kids = None
        return new

    # -- tree ----------------------------------------------------------------------
    @property
    def _children(self):
        if self._kids is None:
            blk = self._block
            self._kids = [CellView._make(blk, i, self) for i in range(len(blk))]
        return self._kids

    @_children.setter
    def _children(self, value):
        self._kids = value

    def _detach_block(self):
        """Turn the block's cells into ordinary Cuboids (keeps the identity of existing views)."""
        blk = self._block
        if blk is None:
            return
        kids = self._children
        self._block = None
        for c in kids:
            d = c.__dict__
            object.__setattr__(c, "__class__", Cuboid)
            i = d.pop("_i")
            d.pop("_blk")
            style = d.pop("_style")
            d["_position"] = blk.pos[i].copy()
            d["_orientation"] = blk.rot
            d["_dimension"] = blk.dim[i].copy()
            d["_polarization"] = blk.pol[i].copy() if blk.pol_set[i] else None
            d["style"] = style if style is not None else Style()
            for k, v in blk.attrs.items():
                d.setdefault(k, v)

    @property
    def children(self):
        return list(self._children)

    def __iter__(self):
        return iter(self._children)

    def __len__(self):
        return len(self._block) if self._kids is None else len(self._kids)

    def __getitem__(self, i):
        return self._children[i]

    def add(self, *children):
        flat = []
        for c in children:
            if isinstance(c, list | tuple):
                flat.extend(c)
            else:
                flat.append(c)
        if flat:
            self._detach_block()
        for c in flat:
            if not isinstance(c, BaseGeo):
                msg = f"cannot add {type(c).__name__!r} to a Collection"
                raise TypeError(msg)
            if c is self:
                msg = "cannot add a Collection to itself"
                raise ValueError(msg)
            if c._parent is self:
                continue  # already a child
            if c._parent is not None:
                c._parent.remove(c)
            c._parent = self
            self._children.append(c)
        return self

    def remove(self, *children):
        if children:
            self._detach_block()
        for c in children:
            self._children = [x for x in self._children if x is not c]
            c._parent = None
        return self

    def _walk(self):
        for c in self._children:
            yield c
            if isinstance(c, Collection):
                yield from c._walk()

    def _source_groups(self):
        """The sources below this collection in ``sources_all`` order, with every un-customised cell
        block as ONE entry (the owning Collection) instead of its n cells."""
        out = []
        if self._block is not None and not self._block.custom:
            out.append(self)
            return out
        for c in self._children:
            if isinstance(c, Collection):
                out.extend(c._source_groups())
            elif isinstance(c, BaseMagnet | BaseCurrent):
                out.append(c)
        return out

    @property
    def children_all(self):
        return list(self._walk())

    @property
    def sources_all(self):
        return [c for c in self._walk() if isinstance(c, BaseMagnet | BaseCurrent)]

    @property
    def sensors_all(self):
        return [c for c in self._walk() if isinstance(c, Sensor)]

    @property
    def collections_all(self):
        return [c for c in self._walk() if isinstance(c, Collection)]

    # -- pose setters move the children along (magpylib Collection semantics) --------
    def _set_position(self, value):
        old = self._position
        if value.ndim == 1 and old.ndim == 1:
            delta = value - old
            if self._block is not None:
                self._block.shift(delta)
            else:
                for c in self._children:
                    c._shift(delta)
        else:
            self._detach_block()
            newp = np.atleast_2d(value)
            oldp = np.atleast_2d(old)
            p = max(len(newp), len(oldp))
            oldp = np.vstack([oldp, np.repeat(oldp[-1:], p - len(oldp), axis=0)])
            newp = np.vstack([newp, np.repeat(newp[-1:], p - len(newp), axis=0)])
            for c in self._children:
                c._shift(newp - oldp)
        self._position = value

    def _set_orientation(self, value):
        rel = value * self._orientation.inv()
        self._rotate_children(rel, self._position)
        self._orientation = value

    def _rotate_children(self, rel, anchor):
        anchor = np.asarray(anchor, dtype=float)
        single = np.asarray(rel.as_quat()).ndim == 1 and anchor.ndim == 1
        if self._block is not None:
            if single:
                self._block.rotate_about(rel, anchor)
                return
            self._detach_block()
        for c in self._children:
            c._rotate_about(rel, anchor)

    def rotate(self, rotation, anchor=None, start="auto"):
        """Rotate the collection AND its children about ``anchor`` (default: the collection's own
        position), as magpylib's Collection does."""
        rot = _as_rotation(rotation)
        anc = self._position if anchor is None else np.array(anchor, dtype=float) * np.ones(3)
        self._rotate_about(rot, anc)
        return self

    def __deepcopy__(self, memo):
        if self._block is not None and self._block.custom:
            self._detach_block()
        if self._block is None:
            return super().__deepcopy__(memo)
        kids, self._kids = self._kids, None  # the views are not copied: the new block makes its own
        try:
            new = super().__deepcopy__(memo)
        finally:
            self._kids = kids
        return new

    def _process_style_kwargs(self, **kwargs):
        out = {}
        for k, v in kwargs.items():
            if not k.startswith("style_"):
                msg = f"unexpected keyword argument {k!r}"
                raise TypeError(msg)
            out[k[6:]] = v
        return out

    def __repr__(self):
        lbl = self.style.label
        return f"Collection(id={self._id}" + (f", label={lbl!r})" if lbl else ")")


def _shift(self, delta):
    """Translate an object (and, for collections, its children) by delta ((3,) or (p,3))."""
    delta = np.asarray(delta, dtype=float)
    if isinstance(self, Collection):
        if self._block is not None and delta.ndim == 1 and self._position.ndim == 1:
            self._block.shift(delta)
        else:
            self._detach_block()
            for c in self._children:
                c._shift(delta)
    if delta.ndim == 2 and self._position.ndim == 1:
        self._pad_orientation(len(delta))
    self._position = self._position + delta


def _rotate_about(self, rel, anchor):
    """Rotate an object (and children) by ``rel`` about ``anchor`` ((3,) or (p,3))."""
    if isinstance(self, Collection):
        self._rotate_children(rel, anchor)
    anchor = np.asarray(anchor, dtype=float)
    relq = np.asarray(rel.as_quat())
    if relq.ndim == 1 and self._position.ndim == 1 and anchor.ndim == 1:
        self._position = anchor + rel.apply(self._position - anchor)
        self._orientation = _rot_single(rel * self._orientation)
        return
    p = max(len(np.atleast_2d(relq)), len(np.atleast_2d(self._position)), len(np.atleast_2d(anchor)))

    def pad(a):
        a = np.atleast_2d(a)
        return np.vstack([a, np.repeat(a[-1:], p - len(a), axis=0)])

    rel_p = R.from_quat(pad(relq))
    pos = pad(self._position)
    anc = pad(anchor)
    self._position = anc + rel_p.apply(pos - anc)
    self._orientation = rel_p * R.from_quat(pad(self._orientation.as_quat()))


BaseGeo._shift = _shift
BaseGeo._rotate_about = _rotate_about


def _merge_observers(observers):
    if len(observers) == 1:
        return observers[0]
    return list(observers)
