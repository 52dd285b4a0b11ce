"""dotted-path access into event bodies (behaviour of mlrun/utils/helpers.py:446-519)."""

_MISSING = object()


def get_in(obj, keys, default=None):
    """`get_in({"a": {"b": 1}}, "a.b") == 1`; a falsy container on the way or a missing key -> default"""
    for key in keys.split(".") if isinstance(keys, str) else keys:
        if not obj or key not in obj:
            return default
        obj = obj[key]
    return obj


def _split(key):
    parts, cur, esc = [], [], False
    for ch in key:
        if ch == "\\":
            esc = not esc
        elif ch == "." and not esc:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts


def update_in(obj, key, value, append=False, replace=True):
    """write `value` at a dotted path (backslash escapes a dot), creating dicts on the way"""
    parts = _split(key) if isinstance(key, str) else list(key)
    for part in parts[:-1]:
        nxt = obj.get(part, _MISSING)
        if nxt is _MISSING:
            nxt = obj[part] = {}
        obj = nxt
    last = parts[-1]
    if last not in obj:
        obj[last] = [] if append else {}
    if append:
        if isinstance(value, list):
            obj[last] += value
        else:
            obj[last].append(value)
    elif replace or not obj.get(last):
        obj[last] = value


def select_input(input_path, body):
    """serving/utils.py:26-31"""
    if not input_path:
        return body
    if not hasattr(body, "__getitem__"):
        raise TypeError("input_path parameter supports only dict-like event bodies")
    return get_in(body, input_path)


def merge_result(result_path, body, result):
    """serving/utils.py:34-43: merge only with a result_path AND a truthy body, else replace"""
    if result_path and body:
        if not hasattr(body, "__getitem__"):
            raise TypeError("result_path parameter supports only dict-like event bodies")
        update_in(body, result_path, result)
        return body
    return result
