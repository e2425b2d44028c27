import json


class JSONIndentEncoder(json.JSONEncoder):
    pass
