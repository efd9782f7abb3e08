"""A second, independent restatement of the reference's physics package — TEST INFRASTRUCTURE ONLY.

Where oracle/ restates the algorithm in C on flat arrays, this file transliterates the Go source line by line and keeps
its data structures: heap objects reached through slices of pointers, the HUD slider table read per particle, and Go's
slice semantics for `removeConstraint` (an in-place `append(s[:i], s[i+1:]...)` while `Cloth.Update` is ranging over the
ORIGINAL slice header).  Python floats are IEEE binary64 with one rounding per operation and `math.sqrt` is correctly
rounded — the arithmetic the Go compiler emits for amd64 — so the two restatements must agree bit for bit; a
transcription slip in either shows up as a difference.  It is still not the reference itself (no Go toolchain in this
image): the oracle stays "parity unpinned".

Only the physics half of (*Cloth).Update is here (physics/cloth.go:92-100); rendering is out of scope.
"""
import math
import struct


def f32(v):
    """float64(float32(v)): what Go's float64(x) of a float32 value yields."""
    return struct.unpack("f", struct.pack("f", v))[0]


# gui/hud.go:32-38 (enum) and :86-92 (table order): the index is the enum value
HudSliderDragForce, HudSliderGravityForce, HudSliderStiffness, HudSliderFriction, HudSliderTearDistance = range(5)
# consts/consts.go:17-20
ClothPinDist, MinFocusArea, MaxFocusArea = 4, 30, 120


class Slider:  # gui/hud.go: a widget.Float (float32 Value) and its float32 Max
    def __init__(self, value, vmax):
        self.Value, self.Max = f32(value), f32(vmax)


class Hud:
    def __init__(self, values, drag_max):
        self.Sliders = {i: Slider(v, drag_max if i == HudSliderDragForce else v) for i, v in enumerate(values)}
        self.WinOffsetX = 0.0
        self.WinOffsetY = 0.0


class Mouse:  # physics/mouse.go:9-19
    def __init__(self):
        self.x = self.y = self.px = self.py = self.force = 0.0
        self.scrollY = 0.0  # unit.Dp = float32
        self.leftDown = self.rightDown = self.isDragging = self.ctrlDown = False


class Particle:  # physics/particle.go:16-27
    __slots__ = ("x", "y", "px", "py", "vx", "vy", "friction", "stiffness", "dragForce", "pinX", "isActive", "highlighted")


def NewParticle(x, y, hud):  # physics/particle.go:30-43
    p = Particle()
    p.x, p.y, p.px, p.py = x, y, x, y
    p.vx = p.vy = p.friction = 0.0
    p.pinX = False
    p.isActive = True
    p.highlighted = False
    p.dragForce = hud.Sliders[HudSliderDragForce].Value
    p.stiffness = hud.Sliders[HudSliderStiffness].Value
    return p


class Constraint:  # physics/constraint.go:10-14
    __slots__ = ("p1", "p2", "length")

    def __init__(self, p1, p2, length):
        self.p1, self.p2, self.length = p1, p2, length


class Cloth:
    def __init__(self, width, height, spacing):  # physics/cloth.go:31-38
        self.Width, self.Height, self.spacing = width, height, spacing
        self.friction = 0.0
        self.particles = []
        # a Go slice = (backing array, len): the array never shrinks, removeConstraint only lowers len
        self.constraints_arr = []
        self.constraints_len = 0
        self.isInitialized = False
        self.colour_of = {}   # not in the Go code: which colour class a stick falls in (UpdateColourOrdered only)
        self.creation = None

    def Init(self, posX, posY, hud):  # physics/cloth.go:42-81
        clothX = self.Width // self.spacing
        clothY = self.Height // self.spacing
        if self.isInitialized:
            return
        for y in range(clothY + 1):
            for x in range(clothX + 1):
                px = posX + x * self.spacing
                py = posY + y * self.spacing
                particle = NewParticle(float(px), float(py), hud)
                particle.friction = self.friction
                if y != 0:
                    top = self.particles[x + (y - 1) * (clothX + 1)]
                    self.constraints_arr.append(Constraint(top, particle, float(self.spacing)))
                    self.colour_of[id(self.constraints_arr[-1])] = 2 + ((y - 1) & 1)  # vertical, head row y-1
                if x != 0:
                    left = self.particles[len(self.particles) - 1]
                    self.constraints_arr.append(Constraint(left, particle, float(self.spacing)))
                    self.colour_of[id(self.constraints_arr[-1])] = (x - 1) & 1        # horizontal, head column x-1
                pinX = x % (clothX // 10)  # ZeroDivisionError where Go panics (clothX < 10)
                if y == 0 and pinX == 0:
                    particle.pinX = True
                self.particles.append(particle)
        self.constraints_len = len(self.constraints_arr)
        self.creation = list(self.constraints_arr)
        self.isInitialized = True

    def Update(self, width, height, mouse, hud, dt):  # physics/cloth.go:85-100 (physics half)
        for p in self.particles:
            particle_update(p, width, height, mouse, hud, dt)
        # `for _, c := range cloth.constraints` evaluates the slice header ONCE: the loop runs over the original length
        # and reads the (possibly shifted) backing array
        arr, n = self.constraints_arr, self.constraints_len
        visited = 0
        for i in range(n):
            c = arr[i]
            if c.p1.isActive and c.p2.isActive:
                constraint_update(c, self, mouse)
                visited += 1
        return visited

    def UpdateColourOrdered(self, width, height, mouse, hud, dt, iterations):
        """BASELINE.json's deterministic mode: the SAME Go functions, fed the colour-ordered stick sequence — H-even, H-odd,
        V-even, V-odd (by the parity of the head particle's column / row), `iterations` times — instead of creation order.
        A stick is visited while it is still in cloth.constraints and both ends are active (cloth.go:97)."""
        for p in self.particles:
            particle_update(p, width, height, mouse, hud, dt)
        visited = 0
        for _ in range(iterations):
            for colour in range(4):
                in_slice = {id(c) for c in self.constraints_arr[:self.constraints_len]}
                for c in self.creation:
                    if self.colour_of[id(c)] == colour and id(c) in in_slice and c.p1.isActive and c.p2.isActive:
                        constraint_update(c, self, mouse)
                        visited += 1
        return visited


def particle_update(p, width, height, mouse, hud, dt):  # physics/particle.go:75-181
    p.highlighted = False
    p.dragForce = hud.Sliders[HudSliderDragForce].Value
    gravityForce = hud.Sliders[HudSliderGravityForce].Value
    p.stiffness = hud.Sliders[HudSliderStiffness].Value
    p.friction = hud.Sliders[HudSliderFriction].Value
    tearDistance = hud.Sliders[HudSliderTearDistance].Value
    if p.pinX:
        p.x += hud.WinOffsetX
        p.y += hud.WinOffsetY
        return
    if mouse.leftDown:
        maxDragForce = hud.Sliders[HudSliderDragForce].Max
        p.dragForce += mouse.force  # increaseForce, :184-189
        if p.dragForce > maxDragForce:
            p.dragForce = maxDragForce
    dx = p.x - mouse.x
    dy = p.y - mouse.y
    dist = math.sqrt(dx * dx + dy * dy)
    if mouse.isDragging and dist < tearDistance:
        dx = mouse.x - mouse.px
        dy = mouse.y - mouse.py
        if dx > p.stiffness:
            dx = p.stiffness
        if dy > p.stiffness:
            dy = p.stiffness
        if dx < -p.stiffness:
            dx = -p.stiffness
        if dy < -p.stiffness:
            dy = -p.stiffness
        p.px = p.x - dx * p.dragForce
        p.py = p.y - dy * p.dragForce
    if mouse.ctrlDown and dist < ClothPinDist:
        p.pinX = True
    focusArea = mouse.scrollY  # float32
    if focusArea > MaxFocusArea:
        focusArea = float(MaxFocusArea)
    elif focusArea < MinFocusArea:
        focusArea = float(MinFocusArea)
    if dist < focusArea:
        p.highlighted = True
    if mouse.rightDown:
        if dist < focusArea:
            p.isActive = False
    px, py = p.x, p.y
    p.vy += gravityForce
    posX, posY = p.vx * (dt * dt), p.vy * (dt * dt)
    p.x = p.x + (p.x - p.px) * p.friction + posX
    p.y = p.y + (p.y - p.py) * p.friction + posY
    p.px, p.py = px, py
    if p.x >= float(width):
        p.x = float(width)
        p.px = p.x
    elif p.x < 0:
        p.x = 0.0
        p.px = p.x
    if p.y > float(height):
        p.y = float(height)
        p.py = p.y
    elif p.y < 0:
        p.y = 0.0
        p.py = p.y
    p.vx, p.vy = 0.0, 0.0


def constraint_update(c, cloth, mouse, tear_length=150, gain=0.35):  # physics/constraint.go:25-54
    dx = c.p1.x - c.p2.x
    dy = c.p1.y - c.p2.y
    dist = math.sqrt(dx * dx + dy * dy)
    if dist < c.length:
        return
    if mouse.isDragging:
        if dist > tear_length:
            remove_constraint(c, cloth)
    diff = (c.length - dist) / dist
    mul = diff * gain * (1 - c.length / dist)
    offsetX, offsetY = dx * mul, dy * mul
    if not c.p1.pinX:
        c.p1.x += offsetX
        c.p1.y += offsetY
    if not c.p2.pinX:
        c.p2.x -= offsetX
        c.p2.y -= offsetY


def remove_constraint(c, cloth):  # physics/constraint.go:57-64
    arr, n = cloth.constraints_arr, cloth.constraints_len
    for idx in range(n):
        if arr[idx] is c:
            # append(s[:idx], s[idx+1:]...): memmove within the same backing array, len - 1; the old last slot keeps
            # its value
            arr[idx:n - 1] = arr[idx + 1:n]
            cloth.constraints_len = n - 1
            break


# ---- the render half's geometry: addSegment / normal (physics/cloth.go:150-168), float32 arithmetic ------------------
# A float32 +, -, *, / of float32 operands equals the binary64 result rounded to float32 (53 >= 2 * 24 + 2 bits).
def go_hypot(p, q):  # math.Hypot: max * sqrt(1 + (min/max)^2) (src/math/hypot.go and hypot_amd64.s), finite arguments
    p, q = abs(p), abs(q)
    if p < q:
        p, q = q, p
    if p == 0:
        return 0.0
    q = q / p
    return p * math.sqrt(1 + q * q)


def normal(a, b, w):  # physics/cloth.go:160-168
    dirx, diry = f32(b[0] - a[0]), f32(b[1] - a[1])
    dirx, diry = +diry, -dirx
    d = go_hypot(dirx, diry)
    if abs(d) < 1e-5:
        return (0.0, 0.0)
    s = f32(w / f32(d))
    return (f32(dirx * s), f32(diry * s))


def add_segment(a, b, w):  # physics/cloth.go:150-157: MoveTo(a+n), LineTo(b+n), LineTo(b-n), LineTo(a-n)
    n = normal(a, b, w)
    return (f32(a[0] + n[0]), f32(a[1] + n[1]), f32(b[0] + n[0]), f32(b[1] + n[1]),
            f32(b[0] - n[0]), f32(b[1] - n[1]), f32(a[0] - n[0]), f32(a[1] - n[1]))
