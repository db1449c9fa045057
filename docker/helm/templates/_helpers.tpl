{{- define "hvd.fullname" -}}{{ .Release.Name }}-hvd{{- end -}}
{{- define "hvd.labels" -}}
app.kubernetes.io/name: horovod-b200
app.kubernetes.io/instance: {{ .Release.Name }}
{{- end -}}
