class Slider:
    pass


class RadioButtons:
    pass


class CheckButtons:
    pass
