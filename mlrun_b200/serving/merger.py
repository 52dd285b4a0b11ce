"""Fan-in join of async graphs: the `Merge` step (plugin-API mirror of mlrun/serving/merger.py:36-156).

Events that reach the step over several branches are joined by event id (default) or by a user key expression.  The
first arrivals of a key wait; the arrival that completes the key releases the joined event.  Only the `max_behind`
(default 64) most recently opened keys may be waiting: opening key number s gives up key number s - max_behind if it is
still incomplete -- it is reported to the error stream and its remaining parts are ignored when they show up.
Subclasses may override `get_join_key` / `merge_function` as with the reference class.

Host-side graph plumbing (no arithmetic).  In the inline executor a waiting arrival ends that branch's walk (`DROP`).
"""

from .step_meta import StepMeta

DROP = object()  # a native step returns it to end the walk of the current event (nothing is emitted downstream)


class _Waiting:
    """the parts of one join key seen so far"""

    __slots__ = ("order", "parts", "given_up")

    def __init__(self, order, first):
        self.order = order        # how many keys had been opened before this one
        self.parts = [first]
        self.given_up = False


class Merge(StepMeta):
    _native_step = True

    def __init__(self, full_event=None, key_path=None, max_behind=None, expected_num_events=None, context=None, name=None,
                 input_path=None, result_path=None, **kwargs):
        self.key_path = key_path
        self.max_behind = max_behind
        self.expected_num_events = expected_num_events
        self._graph_step = kwargs.pop("graph_step", None)
        # joining on event.id (the default) needs the event object, a key expression works on the body
        self._full_event = True if (full_event is None and not key_path) else full_event
        self.context, self.name = context, name
        self._input_path, self._result_path = input_path, result_path
        self._kwargs = kwargs
        self._window = max_behind or 64
        self._fan_in = None
        self._key_of = None
        self._waiting = {}   # join key -> _Waiting (given-up keys stay, so that their late parts are recognised)
        self._opened = {}    # order -> join key, for the keys still inside the window
        self._n_opened = 0

    def post_init(self, mode="sync"):
        branches = len(self._graph_step.after) if self._graph_step else 0
        self._fan_in = self.expected_num_events or branches
        self._key_of = eval("lambda event: " + (self.key_path or "event.id"), {}, {})

    def get_join_key(self, event):
        return self._key_of(event)

    def merge_function(self, last_event, events):
        if not self._full_event:
            return events
        last_event.body = [e.body for e in events]
        return last_event

    # ---- the join ---------------------------------------------------------------------------------
    def _merge_events(self, element):
        """-> the joined event when `element` completes its key, else None"""
        if self._fan_in <= 1:
            return element
        key = self.get_join_key(element)
        ctx = self.context
        slot = self._waiting.get(key)
        if slot is None:
            return self._open(key, element)
        if slot.given_up:
            ctx.logger.warning(f"event id {slot} arrived late")
            return None
        if ctx.verbose:
            ctx.logger.info(f"event id {key} part {len(slot.parts)} arrived")
        slot.parts.append(element)
        if len(slot.parts) < self._fan_in:
            return None
        if ctx.verbose:
            ctx.logger.info(f"event {key}, all {len(slot.parts)} parts arrived")
        del self._waiting[key]
        self._opened.pop(slot.order, None)
        return self.merge_function(element, slot.parts)

    def _open(self, key, element):
        ctx = self.context
        order = self._n_opened
        self._n_opened += 1
        self._waiting[key] = _Waiting(order, element)
        if ctx.verbose:
            ctx.logger.info(f"new event id {key} arrived")
        stale = self._opened.pop(order - self._window, None)
        if stale is not None:  # still incomplete after `max_behind` newer keys: give it up
            message = f"missing parts for event key {stale} after a long wait, dropping"
            ctx.logger.warning(message)
            if self._full_event:
                ctx.push_error(self._waiting[stale].parts[0], f"{message}", source=self.name)
            self._waiting[stale].given_up = True
        self._opened[order] = key
        return None

    def do(self, element):
        """inline-executor entry: the event (full_event) or its body; DROP while parts are missing"""
        joined = self._merge_events(element)
        return joined if joined else DROP
