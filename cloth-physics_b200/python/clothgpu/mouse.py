"""Host-side mirror of physics.Mouse (physics/mouse.go:9-99) and of the event -> Mouse mapping of the
reference's frame loop (main.go:174-177, 260-309), so a scripted trace can be replayed bit for bit
into any implementation of the update.  Pure bookkeeping: no physics here."""
import numpy as np

from . import _abi


def _f32(v):
    return float(np.float32(v))


class Mouse:
    """Same state and accessor names as physics.Mouse (physics/mouse.go:9-19)."""

    def __init__(self):
        self.x = self.y = self.px = self.py = 0.0
        self.force = 0.0
        self.scrollY = 0.0       # unit.Dp (float32)
        self.maxScrollY = 0.0
        self.leftDown = self.rightDown = self.isDragging = self.ctrlDown = False

    # physics/mouse.go:21-27
    def UpdatePosition(self, x, y):
        self.px, self.py = self.x, self.y
        self.x, self.y = x, y

    def SetLeftButton(self): self.leftDown = True
    def GetLeftButton(self): return self.leftDown
    def ReleaseLeftButton(self): self.leftDown = False
    def SetRightButton(self): self.rightDown = True
    def GetRightButton(self): return self.rightDown
    def ReleaseRightButton(self): self.rightDown = False
    def SetDragging(self, d): self.isDragging = bool(d)
    def GetDragging(self): return self.isDragging
    def SetCtrlDown(self, s): self.ctrlDown = bool(s)
    def GetCtrlDown(self): return self.ctrlDown
    def SetForce(self, f): self.force = float(f)
    def GetForce(self): return self.force
    def ResetForce(self): self.force = 0.0
    def SetScrollY(self, s): self.scrollY = _f32(s)
    def GetScrollY(self): return self.scrollY
    def SetMaxScrollY(self, s): self.maxScrollY = _f32(s)
    def GetMaxScrollY(self): return self.maxScrollY

    def fill(self, frame):
        """Copy the Mouse into the fields of a clothgpu_frame (what (*Cloth).Update reads)."""
        frame.mouse_x, frame.mouse_y = self.x, self.y
        frame.mouse_px, frame.mouse_py = self.px, self.py
        frame.mouse_force = self.force
        frame.scroll_y = self.scrollY
        frame.buttons = ((_abi.MOUSE_LEFT if self.leftDown else 0) | (_abi.MOUSE_RIGHT if self.rightDown else 0) |
                         (_abi.MOUSE_DRAGGING if self.isDragging else 0) | (_abi.MOUSE_CTRL if self.ctrlDown else 0))
        return frame


class EventAdapter:
    """The pointer-event handling of main.go:260-309 plus the per-frame force ramp of main.go:174-177,
    with the wall clock replaced by an explicit `seconds_since_press` so traces are deterministic."""

    PRIMARY, SECONDARY = 1, 2

    def __init__(self, mouse=None):
        self.mouse = mouse or Mouse()
        self.mouse.SetScrollY(50.0)        # consts.DefaultFocusArea, main.go:84
        self.mouse.SetMaxScrollY(120.0)    # consts.MaxFocusArea, main.go:85
        self.mouseDrag = False
        self.scroll = 50.0

    def begin_frame(self, seconds_since_press):
        if self.mouse.GetLeftButton():                       # main.go:174-177
            self.mouse.SetForce(seconds_since_press * 5)

    def _buttons(self, x, y, buttons):
        x, y = _f32(x), _f32(y)                              # f32.Point widened, main.go:277
        if buttons == self.PRIMARY:                          # main.go:298-302
            self.mouse.SetLeftButton()
            self.mouse.UpdatePosition(x, y)
            self.mouse.SetDragging(self.mouseDrag)
        elif buttons == self.SECONDARY:                      # main.go:303-306
            self.mouse.SetRightButton()
            self.mouse.UpdatePosition(x, y)

    def scroll_by(self, dy, x=0.0, y=0.0, buttons=0):        # main.go:268-275
        self.scroll = _f32(self.scroll + dy)
        self.scroll = min(max(self.scroll, 30.0), self.mouse.GetMaxScrollY())
        self.mouse.SetScrollY(self.scroll)
        self._buttons(x, y, buttons)

    def move(self, x, y, buttons=0):                         # main.go:276-278
        self.mouse.UpdatePosition(_f32(x), _f32(y))
        self._buttons(x, y, buttons)

    def press(self, x, y, buttons, ctrl=False):              # main.go:279-285
        if ctrl:
            self.mouse.SetCtrlDown(True)
        self.mouse.SetLeftButton()
        self._buttons(x, y, buttons)

    def drag(self, x, y, buttons):                           # main.go:294-295
        self.mouseDrag = True
        self._buttons(x, y, buttons)

    def release(self):                                       # main.go:286-293
        self.mouseDrag = False
        self.mouse.ResetForce()
        self.mouse.ReleaseLeftButton()
        self.mouse.ReleaseRightButton()
        self.mouse.SetDragging(False)
        self.mouse.SetCtrlDown(False)
