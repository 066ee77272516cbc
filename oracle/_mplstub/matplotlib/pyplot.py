"""pyplot stub: every call is a no-op, savefig writes a 2x2 PNG so PIL can open it."""
class _Anything:
    transAxes = None
    def __call__(self, *a, **k):
        return _Anything()
    def __getattr__(self, name):
        return _Anything()
    def __iter__(self):
        return iter(())
    def get_height(self): return 0.0
    def get_x(self): return 0.0
    def get_width(self): return 0.0

def savefig(buf, format="png", **k):
    from PIL import Image
    Image.new("RGB", (2, 2), (255, 255, 255)).save(buf, format="PNG")

def __getattr__(name):
    return _Anything()
