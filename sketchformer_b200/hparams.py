"""Typed `name=value,name=value` hyper-parameter sets: the config surface the reference drives
everything with (utils/hparams.py:232-344 parse rules, 347-697 container, 49-84 helpers).

Written from the behaviour, not the source: a value's type is the type of its default; bools accept
true/True/false/False or an integer; lists are `[a,b,c]` of the element type; an unknown name, a
duplicate name or an unparsable value raises ValueError; ints are accepted for float parameters but
not the other way round.
"""
import json
import re

# One `name=value` clause.  As in the reference's grammar (utils/hparams.py:33-40): the name starts with a letter and
# follows the previous clause's comma (and optional blanks) directly - leading blanks at the very start are malformed;
# blanks around `=` are skipped; a scalar runs up to the next comma or `[` and KEEPS its trailing blanks.
_PAIR = re.compile(r"(?P<name>[A-Za-z][\w.]*)(\[\s*(?P<index>\d+)\s*\])?\s*=\s*"
                   r"((?P<scalar>[^,\[]*)|(?P<list>\[[^\]]*\]))($|,\s*)")


def _parse_scalar(name, typ, text, whole):
    """utils/hparams.py:101-148, 313-328: bools are true / True / false / False or an integer; everything else goes
    through the type's constructor on the text as matched (int() and float() tolerate the trailing blanks, a string
    keeps them and may be empty)."""
    try:
        if typ is bool:
            if text in ("true", "True"):
                return True
            if text in ("false", "False"):
                return False
            return bool(int(text))
        return typ(text)
    except ValueError:
        raise ValueError(f"Could not parse hparam '{name}' of type '{typ.__name__}' with value '{text}' in {whole}")


class HParams:
    """Attribute-style container; `parse("a=1,b=[2,3]")` overrides in place and returns self."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_types", {})
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    # ---- definition ----------------------------------------------------------------------
    def add_hparam(self, name, value):
        if name in self._types or name.startswith("_") or hasattr(type(self), name):
            raise ValueError(f"Hyperparameter name is reserved: {name}")
        if isinstance(value, (list, tuple)):
            if not value:
                raise ValueError(f"Multi-valued hyperparameters cannot be empty: {name}")
            self._types[name] = (type(value[0]), True)
            value = list(value)
        else:
            self._types[name] = (type(value), False)
        object.__setattr__(self, name, value)

    def del_hparam(self, name):
        if name in self._types:
            del self._types[name]
            object.__delattr__(self, name)

    def _cast(self, name, typ, v):
        if isinstance(v, bool) != (typ is bool) and (typ is bool or isinstance(v, bool)):
            raise ValueError(f"Could not cast hparam '{name}' of type '{typ.__name__}' from value {v!r}")
        if typ is int and isinstance(v, float):
            raise ValueError(f"Could not cast hparam '{name}' of type 'int' from value {v!r}")
        if typ in (int, float) and not isinstance(v, (int, float)):
            raise ValueError(f"Could not cast hparam '{name}' of type '{typ.__name__}' from value {v!r}")
        if typ is str and not isinstance(v, str):
            raise ValueError(f"Could not cast hparam '{name}' of type 'str' from value {v!r}")
        return typ(v)

    def set_hparam(self, name, value):
        typ, is_list = self._types[name]          # KeyError for an unknown name, like the reference
        if isinstance(value, (list, tuple)):
            if not is_list:
                raise ValueError(f"Must not pass a list for single-valued parameter: {name}")
            object.__setattr__(self, name, [self._cast(name, typ, v) for v in value])
        else:
            if is_list:
                raise ValueError(f"Must pass a list for multi-valued parameter: {name}.")
            object.__setattr__(self, name, self._cast(name, typ, value))

    def __setattr__(self, name, value):
        if name in self._types:
            self.set_hparam(name, value)
        else:
            object.__setattr__(self, name, value)

    # ---- parsing -------------------------------------------------------------------------
    def parse(self, values):
        out, pos, seen = {}, 0, set()
        values = values or ""
        while pos < len(values):
            m = _PAIR.match(values, pos)
            if not m or m.end() == pos:
                raise ValueError(f"Malformed hyperparameter value: {values[pos:]}")
            pos = m.end()
            name = m.group("name")
            if name not in self._types:
                raise ValueError(f"Unknown hyperparameter type for {name}")
            typ, is_list = self._types[name]
            if m.group("index") is not None:
                if m.group("list") is not None:
                    raise ValueError(f"Could not parse list assigned to an index of '{name}' in {values}")
                idx = int(m.group("index"))
                if (name, idx) in seen or (name, None) in seen:
                    raise ValueError(f"Multiple assignments to variable '{name}[{idx}]' in {values}")
                seen.add((name, idx))
                out.setdefault(name, {})
                if not isinstance(out[name], dict):
                    raise ValueError(f"Multiple assignments to variable '{name}' in {values}")
                out[name][idx] = _parse_scalar(name, typ, m.group("scalar"), values)
                continue
            if (name, None) in seen or name in out:
                raise ValueError(f"Multiple assignments to variable '{name}' in {values}")
            seen.add((name, None))
            if m.group("list") is not None:
                # utils/hparams.py:174-181: elements are separated by commas and / or blanks, empty ones are dropped
                body = m.group("list")[1:-1]
                out[name] = [_parse_scalar(name, typ, t, values) for t in re.split("[ ,]", body) if t]
            else:
                out[name] = _parse_scalar(name, typ, m.group("scalar"), values)
        return self.override_from_dict(out)

    def override_from_dict(self, values):
        # An index assignment `a[1]=x` reaches this point as {1: x}, exactly as in the reference
        # (utils/hparams.py:232-344 builds the dict, 552-567 hands it to set_hparam), where it is then refused:
        # "Must pass a list" for a list hparam, "Could not cast" for a scalar one.  Same here.
        for name, v in values.items():
            self.set_hparam(name, v)
        return self

    # ---- views / (de)serialisation --------------------------------------------------------
    def values(self):
        return {n: getattr(self, n) for n in self._types}

    def get(self, key, default=None):
        return getattr(self, key) if key in self._types else default

    def __contains__(self, key):
        return key in self._types

    def __getitem__(self, key):
        if key not in self._types:
            raise KeyError(key)
        return getattr(self, key)

    def to_json(self, indent=None, separators=None, sort_keys=False):
        return json.dumps(self.values(), indent=indent, separators=separators, sort_keys=sort_keys)

    def parse_json(self, text):
        return self.override_from_dict(json.loads(text))

    def __repr__(self):
        return "HParams(" + ", ".join(f"{k}={v!r}" for k, v in sorted(self.values().items())) + ")"


def copy_hparams(hparams):
    """utils/hparams.py:43-46."""
    return HParams(**hparams.values())


def combine_hparams_into_one(*sets):
    """utils/hparams.py:49-55: one HParams holding every (name, value) of the inputs; later sets win."""
    merged = {}
    for hp in sets:
        merged.update(hp.values())
    return HParams(**merged)


def save_config(output_file, hps, verbose=False):
    """utils/hparams.py:58-67: the merged {model + data} hparams as json (train.py:76-77)."""
    vals = hps.values() if isinstance(hps, HParams) else dict(hps)
    if verbose:
        for k, v in vals.items():
            print(f"{k} = {v}")
    with open(output_file, "w") as f:
        json.dump(vals, f, indent=True, default=lambda o: int(o))
    return vals


def load_config(hps, config_file, verbose=False):
    """utils/hparams.py:70-84: override `hps` from a config.json written by save_config.  Like the
    reference, a failure is reported and leaves `hps` unchanged instead of raising."""
    try:
        with open(config_file) as f:
            hps.parse_json(f.read())
        if verbose:
            for k, v in hps.values().items():
                print(f"{k} = {v}")
    except Exception as e:  # noqa: BLE001 - mirrors the reference's behaviour
        print(f"Error reading config file {config_file}: {e}.\nConfig will not be updated.")
    return hps
